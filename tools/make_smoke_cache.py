"""Write tests/golden/smoke_tiles.packed.npz: the wide lattices of __graft_entry__.smoke(), packed once (the packed
cache format of nfst_b200.data), so that smoke() reaches the tile-stream kernels without running the tensor-op packer.
    python tools/make_smoke_cache.py        (CPU is enough: the packer's layout code is device-agnostic)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import SMOKE_DAG  # noqa: E402
from nfst_b200 import data, synth  # noqa: E402

if __name__ == "__main__":
    ab = synth.random_dag_batch(SMOKE_DAG["B"], SMOKE_DAG["arcs_per_lattice"], levels=SMOKE_DAG["levels"], seed=SMOKE_DAG["seed"])
    p, _ = ab.pack()
    assert p.has_tiles
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "smoke_tiles.packed.npz")
    data.save_packed(out, p, compress=True)
    q = data.load_packed(out)
    assert q.has_tiles and q.n_arcs == p.n_arcs
    print(out, os.path.getsize(out), "bytes;", p.n_arcs, "arcs", [(g.tiles, g.block_threads, g.tile_ring) for g in q.groups])
