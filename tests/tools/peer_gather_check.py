"""Run under torchrun with one process per GPU (>= 2 GPUs): every rank's kernels write their shard of the results
straight into rank 0's HBM over NVLink (shard.PeerGather); rank 0 compares the gathered arrays with the golden
reference answers. Prints PEER_GATHER_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bloomfiltertrie_b200 import engine, shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")          # handles and barriers only; the data path is the kernels' own peer stores
g = os.path.join(ROOT, "tests", "golden")
ok = True
for name in ("golden_pan_k27_g100", "golden_deep_k63_g12", "golden_canon_k27_g8"):
    z = np.load(os.path.join(g, name + ".npz"))
    eng = engine.BFTEngine(os.path.join(g, name + ".bft"), device=local)
    pg = shard.PeerGather(eng)
    reps = 40                              # enough sequences / k-mers for every rank to own a real slice
    offs = z["seq_offs"].astype(np.uint64)
    chars = np.tile(z["seq_chars"], reps)
    all_offs = np.concatenate([[0], np.cumsum(np.tile(np.diff(offs), reps))]).astype(np.uint64)
    rows = pg.query_sequences(chars, all_offs, 0.8, False)
    q = np.tile(z["queries"], (reps, 1))
    present, krows = pg.query_kmers(q)
    if rank == 0:
        ok &= np.array_equal(rows, np.tile(z["seqrows_c0_t0.8"], (reps, 1)))
        ok &= np.array_equal(present, np.tile(z["present"], reps))
        ok &= np.array_equal(krows, np.tile(z["rows"], (reps, 1)))
    eng.close()
dist.barrier()
if rank == 0:
    print("PEER_GATHER_OK" if ok else "PEER_GATHER_MISMATCH", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
