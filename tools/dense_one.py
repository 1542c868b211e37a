import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hubbardtn_b200 import device, sectors as S, synthetic
ctx = device.Context(0)
n, chi = 1024, 6
case = synthetic.HeffCase(ctx, S.U1U1, D=n, chi=chi, spaces=({(0, 0, 0): n}, {(0, 0, 0): n}, [(0, 0, 0)], [(0, 0, 0)] * chi))
for _ in range(3):
    case.plan.apply(case.x, case.y)
ctx.synchronize()
print(case.plan.profile(case.x, case.y, reps=3))
