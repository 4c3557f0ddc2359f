"""world_size-2 gloo worker for tests/test_host_logic.py (CPU only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))

import torch
import torch.distributed as dist


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcabo_b200.parallel import shard_runs, candidate_range, global_argmax
    runs = [(f, i) for f in (15, 16, 17) for i in range(30)]
    mine = shard_runs(runs, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat = [tuple(x) for part in gathered for x in part]
    assert sorted(flat) == sorted(runs) and len(set(flat)) == len(runs)
    # candidate-sharded argmax with ties across ranks -> lowest global index wins (torch.argmax semantics)
    N = 1000
    g = torch.Generator()
    g.manual_seed(0)
    vals = torch.randn(N, dtype=torch.float64, generator=g)
    vals[[123, 777]] = vals.max() + 1.0
    lo, hi = candidate_range(N, rank, world)
    local = vals[lo:hi]
    li = int(torch.argmax(local))
    bv, bi = global_argmax(local[li].reshape(1), torch.tensor([lo + li]), group=None)
    assert int(bi) == 123 and float(bv) == float(vals[123]), (bv, bi)
    if rank == 0:
        print("argmax ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
