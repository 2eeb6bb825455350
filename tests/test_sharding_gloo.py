"""The N > 1 path on CPU: two ranks (gloo) plan shards of one batch with the C helper, each encodes
its shard with the CPU checker standing in for a GPU, and the host-side offset concatenation rebuilds
exactly what a single process produces. No collective touches payload data."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as graft
    import refcodec
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = graft.load_package()
    lib = pkg.product_library()
    oracle = refcodec.OracleLib()
    table = oracle.table(*refcodec.table_arrays("hpack"))

    rng = np.random.default_rng(123)  # same batch on every rank
    data, offs = refcodec.random_batch(rng, 5000, 0, 300, "hpack")
    begin = lib.plan_shards(offs, world)
    a, b = int(begin[rank]), int(begin[rank + 1])
    local_offs = (offs[a:b + 1] - offs[a]).astype(np.uint64)
    local_data = data[int(offs[a]):int(offs[b])]
    enc = oracle.encode_batch(table, 0xFF, local_data, local_offs, 4 * len(local_data) + 16)
    total = int(enc["out_offsets"][-1])

    # every rank publishes its shard-local offsets + payload size; rank 0 concatenates on the host
    gathered = [None] * world
    dist.all_gather_object(gathered, (enc["out_offsets"], enc["out"][:total].tobytes()))
    if rank == 0:
        glob_offs = lib.concat_offsets([g[0] for g in gathered])
        payload = b"".join(g[1] for g in gathered)
        whole = oracle.encode_batch(table, 0xFF, data, offs, 4 * len(data) + 16)
        assert np.array_equal(glob_offs, whole["out_offsets"])
        assert payload == whole["out"][:int(whole["out_offsets"][-1])].tobytes()
        sizes = [int(offs[begin[r + 1]] - offs[begin[r]]) for r in range(world)]
        assert max(sizes) - min(sizes) <= 600, sizes
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_concat(pkg, tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
