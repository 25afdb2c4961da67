"""First-light diagnostics on a B200: GEMM building block, radius graph, forward vs oracle (per-block trace)."""
import os, sys, time, traceback
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_npz_groups
from diffndm_b200 import engine as E
from diffndm_b200.weights import DynamicsConfig, random_init
from oracle import egnn_oracle as O

dev = torch.device('cuda')
print(torch.cuda.get_device_name(0), E.load_library().dndm_version().decode(), flush=True)

def gemm_case(M, N, K, act=0, bias=True, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev) if bias else None
    out = E.test_gemm(a, w, b, act)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    if bias: ref = ref + b
    if act: ref = torch.nn.functional.silu(ref)
    err = (out - ref).abs()
    print(f'gemm M={M} N={N} K={K} act={act}: max_err={err.max().item():.3e} ref_max={ref.abs().max().item():.3f}', flush=True)
    if err.max().item() > 1e-2:
        bad = (err > 1e-2)
        print('  bad frac', bad.float().mean().item(), 'bad rows', bad.any(1).sum().item(), 'bad cols', bad.any(0).sum().item())
        print('  first bad idx', bad.nonzero()[:8].tolist())
        print('  out[0,:8]', out[0, :8].tolist(), '\n  ref[0,:8]', ref[0, :8].tolist())
        r = out / ref
        print('  ratio median', r.median().item())
    return err.max().item()

def run(name, fn):
    try:
        fn()
    except Exception:
        print(f'!! {name} raised'); traceback.print_exc()
    sys.stdout.flush()

def t_gemm():
    gemm_case(128, 128, 64, bias=False)
    gemm_case(128, 128, 256)
    gemm_case(300, 512, 256, act=1)
    gemm_case(1000, 256, 512)
    gemm_case(77, 1024, 256)
run('gemm', t_gemm)

cfg = DynamicsConfig()
W = random_init(cfg, 0, 0.3)
eng = E.Engine(cfg, max_nodes=4096, max_edges=200000, max_samples=16)
eng.load_weights(W)

def t_graph():
    cases, _ = load_npz_groups('edges.npz')
    for name, c in sorted(cases.items()):
        B = int(c['lig_mask'].max()) + 1
        rp, col = eng.radius_graph(torch.from_numpy(c['xh_lig']).to(dev), torch.from_numpy(c['xh_pocket']).to(dev),
                                   torch.from_numpy(c['lig_mask']).to(dev), torch.from_numpy(c['pocket_mask']).to(dev), B)
        rp, col = rp.cpu().numpy(), col.cpu().numpy()
        ref = c['edges']
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        ok = len(col) == ref.shape[1] and np.array_equal(rows, ref[0]) and np.array_equal(col, ref[1])
        print(f'graph {name}: E={len(col)} ref={ref.shape[1]} exact={ok}', flush=True)
run('graph', t_graph)

def t_forward():
    cases, _ = load_npz_groups('forward.npz')
    for name, c in sorted(cases.items()):
        B = len(c['t'])
        N = len(c['lig_mask']) + len(c['pocket_mask'])
        n_l = len(c['lig_mask'])
        trh, trx = eng.set_trace(N)
        ol, op = eng.forward(torch.from_numpy(c['xh_lig']).to(dev), torch.from_numpy(c['xh_pocket']).to(dev),
                             torch.from_numpy(c['t']).to(dev), torch.from_numpy(c['lig_mask']).to(dev),
                             torch.from_numpy(c['pocket_mask']).to(dev), B)
        torch.cuda.synchronize()
        flags = eng.read_flags()
        e_, el_ = eng.graph_stats()
        ol, op = ol.cpu().numpy(), op.cpu().numpy()
        trace = {}
        O.dynamics_forward(W, c['xh_lig'], c['xh_pocket'], c['t'], c['lig_mask'], c['pocket_mask'], O.OracleConfig(),
                           dtype=np.float64, trace=trace)
        print(f'forward {name}: flags={flags} E={e_} E_lig={el_} (oracle E={trace["edges"].shape[1]})')
        for i in range(cfg.n_layers):
            h = trh[i, :N].cpu().numpy(); x = trx[i, :N].cpu().numpy()
            dh = np.abs(h - trace[f'h_{i}']); dx = np.abs(x - trace[f'x_{i}'])
            print(f'   block {i}: h err max={dh.max():.3e} (lig {dh[:n_l].max():.3e}) rel={dh.max()/np.abs(trace[f"h_{i}"]).max():.3e} '
                  f'| x err max={dx.max():.3e} nan_h={np.isnan(h).sum()} nan_x={np.isnan(x).sum()}')
        ex = np.abs(ol[:, :3] - c['out_lig_f64'][:, :3]).max(); eh = np.abs(ol[:, 3:] - c['out_lig_f64'][:, 3:]).max()
        ep = np.abs(op - c['out_pocket_f64']).max()
        print(f'   out: eps_x err={ex:.3e} (max {np.abs(c["out_lig_f64"][:, :3]).max():.3f}) eps_h err={eh:.3e} '
              f'(max {np.abs(c["out_lig_f64"][:, 3:]).max():.3f}) pocket err={ep:.3e}', flush=True)
run('forward', t_forward)
