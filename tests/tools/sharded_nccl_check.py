"""Run under torchrun with one process per GPU (>= 2 GPUs), backend NCCL: shard.ShardedQuery splits the golden query
batches over the ranks, every rank answers its shard on its own GPU, results are gathered over NCCL and rank 0 compares
them with the golden reference answers; then the fused reduction of the sharded k-mer path — every rank's query kernel
adding its hit count into ONE counter in rank 0's HBM through a CUDA-IPC mapping (bft_b200_query_kmers_device_accumulate)
— is checked against the golden number of present k-mers. Prints SHARDED_NCCL_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bloomfiltertrie_b200 import engine, shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = os.path.join(ROOT, "tests", "golden")
ok = True
why = []
for name in ("golden_pan_k27_g100", "golden_deep_k63_g12", "golden_canon_k27_g8"):
    z = np.load(os.path.join(g, name + ".npz"))
    eng = engine.BFTEngine(os.path.join(g, name + ".bft"), device=local)
    sq = shard.ShardedQuery(eng)
    reps = 16
    q = np.tile(z["queries"], (reps, 1))
    present, rows = sq.query_kmers(q)
    offs = z["seq_offs"].astype(np.uint64)
    chars = np.tile(z["seq_chars"], reps)
    all_offs = np.concatenate([[0], np.cumsum(np.tile(np.diff(offs), reps))]).astype(np.uint64)
    srows = sq.query_sequences(chars, all_offs, 0.8, False)
    n_br = sq.query_branching_count(q)
    want_br = reps * int(((z["succ"] > 1) | (z["pred"] > 1)).sum())
    if n_br != want_br:
        ok = False
        why.append(f"{name}: branching {n_br} != {want_br}")
    if rank == 0:
        for label, got, want in (("present", present, np.tile(z["present"], reps)), ("rows", rows, np.tile(z["rows"], (reps, 1))),
                                 ("seqrows", srows, np.tile(z["seqrows_c0_t0.8"], (reps, 1)))):
            if not np.array_equal(got, want):
                ok = False
                why.append(f"{name}: {label} differ")
    # fused reduction: one counter on rank 0, every rank's kernel adds its shard's hits over NVLink
    if eng.RW in (1, 2, 4):
        base = eng.device_alloc(8) if rank == 0 else 0
        box = [eng.peer_export(base) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ptr = base if rank == 0 else eng.peer_import(box[0])
        b, e = shard.shard_range(len(q), rank, world)
        d_q = torch.from_numpy(np.ascontiguousarray(q[b:e]).view(np.int64)).to(dev)
        d_p = torch.empty(e - b, dtype=torch.uint8, device=dev)
        d_r = torch.empty((e - b, eng.RW), dtype=torch.int32, device=dev)
        for _ in range(3):
            eng.query_kmers_device_accumulate(d_q, e - b, d_p, d_r, ptr)
        eng.sync()
        dist.barrier()
        if rank == 0:
            total = int(eng.copy_from_device(base, np.zeros(1, dtype=np.uint64))[0])
            want = 3 * reps * int(z["present"].sum())
            if total != want:
                ok = False
                why.append(f"{name}: peer counter {total} != {want}")
        dist.barrier()
        if rank != 0:
            eng.peer_close(ptr)
        dist.barrier()
        if rank == 0:
            eng.device_free(base)
    eng.close()
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
ok = int(flag.item()) == 0
if rank == 0:
    print("SHARDED_NCCL_OK" if ok else "SHARDED_NCCL_MISMATCH " + "; ".join(why), f"world={world}", flush=True)
elif why:
    print(f"rank {rank}: " + "; ".join(why), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
