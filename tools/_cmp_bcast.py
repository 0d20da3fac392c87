import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import basix_b200 as bx
def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r
for fam, deg, name in ((3, 3, "N1curl3"), (2, 4, "RT4")):
    e = bx.create_element(fam, 3, deg, 11)
    nc = 8_000_000
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    J = torch.eye(3, device="cuda", dtype=torch.float64)[None] + 0.3 * torch.randn(nc, 3, 3, device="cuda", dtype=torch.float64, generator=g)
    detJ = torch.linalg.det(J); K = torch.linalg.inv(J).contiguous()
    X = torch.tensor([[0.2, 0.3, 0.1]], device="cuda", dtype=torch.float64)
    U = e.tabulate(0, X)[0]  # (1, ndofs, 3)
    U2 = U[0].contiguous()
    for rep in range(2):
        ms2, r2 = t(lambda: e.physical_basis(None, J=J, detJ=detJ, K=K, U=U))
        ms1, r1 = t(lambda: e.push_forward_broadcast(U2, J, detJ, K))
        print(rep, ms1, ms2)
    print(name, "bcast ms", ms1, "fused(flat) ms", ms2, "equal", bool(torch.equal(r1.reshape(-1), r2.reshape(-1))), r1.shape, r2.shape)
