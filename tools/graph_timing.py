import sys, torch
sys.path.insert(0, '/root/repo')
import nfst_b200 as nb
from nfst_b200 import synth
from nfst_b200.pack import concat_packed
dev = torch.device('cuda', 0)
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cases = [("config1 B=32", lambda: synth.transliteration_batch(32, seed=0), 50),
         ("SNIPS B=256", lambda: synth.snips_batch(256, seed=1), 50),
         ("dag 300k B=64", lambda: synth.random_dag_batch(64, 300_000, seed=3, device=dev), 10),
         ("dag 1M B=32", lambda: synth.random_dag_batch(32, 1_000_000, seed=3, device=dev), 10)]
for name, gen, reps in cases:
    p, sc = gen().to(dev).pack()
    fb = lambda: nb.lattice_forward_backward(p, arc_scores=sc)
    t_eager = timed(fb, reps)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fb()
    t_graph = timed(g.replay, reps)
    print(f"{name}: eager {t_eager*1e3:.1f} us, graph replay {t_graph*1e3:.1f} us, arcs {p.n_arcs}, launches/step ~{sum(2*(g_.n_levels or 1) for g_ in p.groups)}")
