"""CPU oracle for the HubbardTN hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

This package is a plain numpy restatement of the arithmetic that
DaanVrancken/HubbardTN delegates to MPSKit 0.13.1 / TensorKit 0.14.6 /
KrylovKit 0.9.5 (pinned in /root/reference/Manifest.toml:548,722,1156): block
sparse U(1)xSU(2) / U(1)xU(1) symmetric tensors, the effective-Hamiltonian
applications (H_AC, H_C, H_AC2), the environment transfer updates, the Lanczos /
GMRES solvers, positive QR gauge fixing, per-sector truncated SVD and the IDMRG2 /
VUMPS drivers called at src/HubbardFunctions.jl:1010-1027.

Those packages are NOT vendored in /root/reference and Julia is not installed, so
the algorithms are restated from their published form and the
reference's own call sites, operator definitions (HubbardFunctions.jl:245-472) and
golden energies (test/OB.jl:21,44; test/Spin.jl:42; test/MB.jl:59).

Pinning status (see DESIGN.md "Oracle"):
  * every block-sparse contraction is checked against its dense (symmetry-free)
    expansion, and the MPO against an exact-diagonalisation Hamiltonian;
  * end-to-end energies are checked against the reference's golden vectors and the
    Lieb-Wu Bethe-ansatz values: the full HF:993-1030 schedule reproduces the printed
    digits of test/OB.jl:21,44 (one-band, 2- and 4-site cells) and of test/MB.jl:59
    (-0.630375296, two-band 4-site cell, which also pins that IDMRG2 starts from the
    infinite environments of the initial state as MPSKit does);
  * intermediate quantities (block tables, single applies) are unpinned by the
    reference (it has no such tests): parity for those is "oracle-defined".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (hubbardtn_b200) never does.
"""
