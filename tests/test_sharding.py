"""CPU: utterance sharding (LPT partition, length buckets) and the world_size-2 gloo exchange of output
lengths / metrics -- the only collectives of the multi-GPU path."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ims_toucan_prosody_variance_b200 import sharding


def test_partition_lpt_properties():
    rng = random.Random(0)
    lens = [rng.randint(20, 200) for _ in range(512)]
    costs = [sharding.estimate_cost(n) for n in lens]
    for world in (1, 2, 4, 8):
        shards = sharding.partition_lpt(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(512))          # every utterance exactly once
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= min(loads) + max(costs)                              # LPT bound
        assert max(loads) / (sum(loads) / world) < 1.02                           # near-linear scaling at 512 utterances
        for s in shards:
            assert [costs[i] for i in s] == sorted((costs[i] for i in s), reverse=True)
    assert sharding.partition_lpt([], 4) == [[], [], [], []]
    assert sharding.partition_lpt([3.0], 2) == [[0], []]


def test_bucket_by_length_bounds_padding():
    rng = random.Random(1)
    lens = [rng.randint(20, 200) for _ in range(100)]
    batches = sharding.bucket_by_length(range(100), lens, max_batch=16, max_padding_ratio=1.25)
    assert sorted(i for b in batches for i in b) == list(range(100))
    for b in batches:
        assert len(b) <= 16
        assert max(lens[i] for i in b) <= 1.25 * min(lens[i] for i in b)
    assert sharding.bucket_by_length([], lens, 4) == []
    # default policy: no ratio bound -> length-sorted batches, all full except the last
    batches = sharding.bucket_by_length(range(100), lens, max_batch=16)
    assert [len(b) for b in batches] == [16] * 6 + [4]
    flat = [lens[i] for b in batches for i in b]
    assert flat == sorted(lens, reverse=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [20 + (13 * i) % 181 for i in range(n_total)]
    shards = sharding.partition_lpt([sharding.estimate_cost(n) for n in lens], world)
    mine = shards[rank]
    out_len = [lens[i] * 5 * 384 for i in mine]                                  # what this rank "synthesised"
    all_len = sharding.gather_output_lengths(mine, out_len, n_total)
    audio, ms = sharding.reduce_metrics(sum(out_len) / 24000.0, 10.0 + rank)
    torch.save({"lengths": all_len, "audio": audio, "ms": ms, "mine": mine}, os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gloo_world2_length_gather(tmp_path):
    world, n_total = 2, 37
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    lens = [20 + (13 * i) % 181 for i in range(n_total)]
    expect = torch.tensor([n * 5 * 384 for n in lens], dtype=torch.int64)
    for r in res:
        assert torch.equal(r["lengths"], expect)                                 # every rank sees every length
        assert abs(r["audio"] - float(expect.sum()) / 24000.0) < 1e-6            # SUM over ranks
        assert r["ms"] == 11.0                                                   # MAX over ranks
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(n_total))


def test_single_process_is_identity():
    out = sharding.gather_output_lengths([2, 0], [30, 10], 3)
    assert out.tolist() == [10, 0, 30]
    assert sharding.reduce_metrics(1.5, 2.5) == (1.5, 2.5)
