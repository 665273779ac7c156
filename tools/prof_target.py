"""Target for ncu captures: N rhs! calls of 3D Euler EC p=3 at the given level (default 5)."""
import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
sys.path[:0] = [os.path.join(sys.path[0], "tests"), os.path.join(sys.path[0], "oracle")]
import torch
import cases
lv = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
name = sys.argv[3] if len(sys.argv) > 3 else "c5_euler_ec_3d"
staged = len(sys.argv) > 4 and sys.argv[4] == "staged"
c = dict(cases.CASES[name], level=lv)
semi = cases.make_semi(c, node_coordinates=False, staged_only=staged)
u = semi.compute_coefficients_gpu(0.0, on_device=True)
du = semi.new_vector()
for _ in range(n):
    semi.rhs(du, u, 0.0)
torch.cuda.synchronize()
print("done", semi.nelements, semi.launch_count())
