"""-m gpu, needs >= 2 devices: pool scoring sharded over ranks (frame i -> rank i mod W) with ONE all-gather of the
per-frame records gives every rank the same dictionary as a single-GPU run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crb3d import scorer, second, synth
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    model = second.SECONDNet().eval().to_device(dev)
    frames = [synth.make_frame(50 + i)[::2] for i in range(7)]
    ps = scorer.PoolScorer(model, dev, batch_size=2)
    first = ps.to_device(ps.stage_host(frames[:2]))
    second.calibrate_head_bias(model, first[0], first[1], 2, target_fraction=0.004)
    recs = ps.score_pool(frames)
    q.put((rank, {k: (round(v["entropy"], 6), v["labels"].tolist(), np.round(v["density"], 4).tolist()) for k, v in recs.items()}))
    dist.destroy_process_group()


def test_score_pool_two_gpus_matches_one(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from crb3d import scorer, second, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=600) for _ in procs)
    [p.join(timeout=120) for p in procs]
    assert res[0] == res[1] and sorted(res[0]) == list(range(7))
    torch.manual_seed(0)
    model = second.SECONDNet().eval().to_device(cuda)
    frames = [synth.make_frame(50 + i)[::2] for i in range(7)]
    ps = scorer.PoolScorer(model, cuda, batch_size=2)
    first = ps.to_device(ps.stage_host(frames[:2]))
    second.calibrate_head_bias(model, first[0], first[1], 2, target_fraction=0.004)
    single = ps.score_pool(frames)
    # cuDNN may pick different TF32 algorithms in different processes: near-tied candidate scores can then swap, so the
    # comparison with the single-process run is statistical; the sharding/gather logic itself is checked exactly above
    for k in range(7):
        n_multi, n_single = len(res[0][k][1]), len(single[k]["labels"])
        assert abs(n_multi - n_single) <= max(3, n_single // 10), (k, n_multi, n_single)
        assert abs(res[0][k][0] - single[k]["entropy"]) < 0.1, (k, res[0][k][0], single[k]["entropy"])
