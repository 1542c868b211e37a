"""Run the grouped GEMM kernel on uniform dense problems (one symmetry sector) to separate the
kernel's inner-loop efficiency from the ragged-block / short-K effects of the real workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hubbardtn_b200 import device, sectors as S, synthetic

ctx = device.Context(0)
peak = ctx.probe_fp64_peak(0)
print("DMMA peak %.2f TF/s" % peak)
for n, chi in [(2048, 7), (1024, 13), (512, 25), (192, 97), (167, 97), (128, 97)]:
    spaces = ({(0, 0, 0): n}, {(0, 0, 0): n}, [(0, 0, 0)], [(0, 0, 0)] * chi)
    case = synthetic.HeffCase(ctx, S.U1U1, D=n, chi=chi, spaces=spaces)
    st = case.plan.stats
    pr = case.plan.profile(case.x, case.y, reps=5)
    print("n=%4d chi=%3d  L: %7.1f us %5.2f TF/s (%4.1f%%)   R: %7.1f us %5.2f TF/s (%4.1f%%)  W %6.1f us  tilesL %d tilesR %d"
          % (n, chi, pr["stage_L_ms"] * 1e3, st["flops_L"] / pr["stage_L_ms"] / 1e9, 100 * st["flops_L"] / pr["stage_L_ms"] / 1e9 / peak,
             pr["stage_R_ms"] * 1e3, st["flops_R"] / pr["stage_R_ms"] / 1e9, 100 * st["flops_R"] / pr["stage_R_ms"] / 1e9 / peak,
             pr["stage_W_ms"] * 1e3, st["n_tiles_L"], st["n_tiles_R"]))
