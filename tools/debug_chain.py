"""Developer script: per-16-column error map of the fused chain kernel against fp64 (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from nero_b200 import ops
from nero_b200.ops import Mat, chain, chain_layer as CL
from test_gemm_gpu import _mk_layer
dev = torch.device('cuda')
sp = lambda x: torch.nn.functional.softplus(x, beta=100)

def blocks(got, want, name):
    e = (got.double() - want).abs()
    per = [float(e[:, c:c + 16].max()) for c in range(0, e.shape[1], 16)]
    rows = e.max(dim=1)[0]
    bad = torch.nonzero(rows > 1e-3)[:, 0]
    print(f'{name}: max {float(e.max()):.2e}; per 16-col block:', ' '.join(f'{x:.0e}' for x in per))
    if bad.numel():
        print(f'   bad rows: {bad.numel()} of {e.shape[0]}, first {bad[:8].tolist()} last {bad[-4:].tolist()}')

M = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for (K, N, act) in [(256, 256, 'sp'), (256, 224, 'sp'), (256, 217, 'sp'), (256, 128, 'sp'), (256, 64, 'sp'), (64, 256, 'sp')]:
    L, W, b = _mk_layer(ops, N, K, dev, seed=K + N)
    X = torch.randn(M, 256, device=dev) * 0.3
    out = torch.zeros(M, 256, device=dev)
    chain(Mat(X), K, [CL(L, ops.EK_BIAS_SOFTPLUS, N, save=Mat(out), write_a=False)])
    torch.cuda.synchronize()
    want = sp(X[:, :K].double() @ W.t() + b)
    blocks(out[:, :N], want, f'single layer K={K} N={N}')
# two layers
L0, W0, b0 = _mk_layer(ops, 256, 256, dev, seed=1)
for N1 in (256, 224, 217):
    L1, W1, b1 = _mk_layer(ops, N1, 256, dev, seed=2)
    X = torch.randn(M, 256, device=dev) * 0.3
    H1, H2 = torch.zeros(M, 256, device=dev), torch.zeros(M, 256, device=dev)
    chain(Mat(X), 256, [CL(L0, ops.EK_BIAS_SOFTPLUS, 256, save=Mat(H1)), CL(L1, ops.EK_BIAS_SOFTPLUS, N1, save=Mat(H2), write_a=False)])
    torch.cuda.synchronize()
    h1 = sp(X.double() @ W0.t() + b0)
    h2 = sp(h1 @ W1.t() + b1)
    blocks(H1, h1, f'two layers N1={N1}: h1')
    blocks(H2[:, :N1], h2, f'two layers N1={N1}: h2')

print('---- the 4-layer test chain')
L0, W0, b0 = _mk_layer(ops, 256, 39, dev, seed=1, t_cols=(0, 39))
L1, W1, b1 = _mk_layer(ops, 217, 256, dev, seed=2, t_cols=(0, 256))
L2, W2, b2 = _mk_layer(ops, 256, 256, dev, seed=3, t_cols=(0, 256))
L3, W3, b3 = _mk_layer(ops, 3, 256, dev, seed=4, t_cols=(0, 256))
for variant in ('full', 'no_csrc', 'two_layers_only', 'no_oscale'):
    X = torch.zeros(M, 64, device=dev)
    X[:, :39] = torch.randn(M, 39, device=dev) * 0.5
    H1, H2, H3 = torch.zeros(M, 256, device=dev), torch.zeros(M, 256, device=dev), torch.zeros(M, 256, device=dev)
    H2[:, 217:] = X[:, :39] * 0.5
    out = torch.zeros(M, 32, device=dev)
    osc = 1.0 if variant == 'no_oscale' else 0.70710678
    cs = None if variant == 'no_csrc' else Mat(H2)
    ls = [CL(L0, ops.EK_BIAS_SOFTPLUS, 256, save=Mat(H1)), CL(L1, ops.EK_BIAS_SOFTPLUS, 217, oscale=osc, save=Mat(H2), csrc=cs)]
    if variant != 'two_layers_only':
        ls += [CL(L2, ops.EK_BIAS_SOFTPLUS, 256, save=Mat(H3)), CL(L3, ops.EK_BIAS_GENERIC, 3, act=3, save=Mat(out, 8), write_a=False)]
    chain(Mat(X), 40, ls)
    torch.cuda.synchronize()
    h1 = sp(X[:, :39].double() @ W0.t() + b0)
    tail = X[:, :39].double() * 0.5
    h2 = torch.cat([osc * sp(h1 @ W1.t() + b1), tail], -1)
    h2in = h2 if cs is not None else torch.cat([h2[:, :217], torch.zeros(M, 39, device=dev, dtype=torch.float64)], -1)
    h3 = sp(h2in @ W2.t() + b2)
    o = torch.sigmoid(h3 @ W3.t() + b3)
    print('variant', variant)
    blocks(H1, h1, '  h1')
    blocks(H2, h2, '  h2 (incl. tail)')
    if variant != 'two_layers_only':
        blocks(H3, h3, '  h3')
        blocks(out[:, 8:11], o, '  out')
