"""The exchange step of the path on real hardware: two ranks, one GPU each, detections and
target statistics stored into the peer's buffer over NVLink (dspnet_b200.dist.P2PDetectionGatherer).  One rank is
deliberately slow, and in the second mode the fast rank does not read most generations: the flow control (sequence
numbers + acknowledgements) must make it wait instead of overwriting a slot its peer has not read, and what is read
must always be the rows of THAT step (every step has its own inputs).  Needs two GPUs (gpurun --gpus 2); skipped on
the one-GPU box."""
import os
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util

pytestmark = pytest.mark.gpu
STEPS, BPR, K = 9, 2, 200


def _expected(oracle, step):
    anchors, prob, lp = util.detection_inputs(oracle, "ssd300", 2 * BPR, config_id=50, first_image=100 * step)
    want = oracle.multibox_detection(prob, lp, anchors, nms_threshold=0.45, nms_topk=400)
    rows = np.full((2 * BPR, K, 7), -1.0, np.float32)
    counts = np.zeros((2 * BPR,), np.int32)
    for b in range(2 * BPR):
        keep = want[b][want[b, :, 0] >= 0][:K]
        rows[b, : len(keep)] = keep
        counts[b] = len(keep)
    return anchors, prob, lp, rows, counts


def _worker(rank, port, consume_every, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from oracle import oracle as O
    from dspnet_b200 import MultiBoxDetection
    from dspnet_b200.dist import P2PDetectionGatherer
    ok, g = True, None
    try:
        g = P2PDetectionGatherer(BPR, 8732, K, dev, 2, rank, stats_width=4)
        outs = []
        for step in range(STEPS):
            anchors, prob, lp, rows, counts = _expected(O, step)
            sl = slice(rank * BPR, (rank + 1) * BPR)
            t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
            out = MultiBoxDetection(t(prob[sl]), t(lp[sl]), t(anchors), nms_threshold=0.45, nms_topk=400)
            stats = torch.full((BPR, 4), 1000 * step + 10 * rank, dtype=torch.int32, device=dev)
            outs.append((out, stats))  # submit() reads them on a side stream: keep them alive and untouched
            if rank == 1:
                time.sleep(0.05)       # the slow rank: the other one runs up to its flow-control limit
            g.submit(out, step, stats=stats)
            if step % consume_every == consume_every - 1 or step == STEPS - 1:
                got_rows, got_counts = g.gathered(step)
                got_stats = g.gathered_stats(step)
                torch.cuda.synchronize(dev)
                ok &= bool(np.array_equal(got_counts.cpu().numpy(), counts))
                ok &= bool(np.array_equal(got_rows.cpu().numpy().view(np.uint32), rows.view(np.uint32)))
                want_stats = np.repeat(np.array([1000 * step, 1000 * step + 10], np.int32), BPR)[:, None].repeat(4, 1)
                ok &= bool(np.array_equal(got_stats.cpu().numpy(), want_stats))
        ok &= g.check()
    except Exception as e:  # noqa: BLE001 -- reported through the result file
        ok = False
        print("rank %d: %r" % (rank, e), flush=True)
    finally:
        if g is not None:
            g.close()
        dist.destroy_process_group()
    with open(result + ".%d" % rank, "w") as f:
        f.write("ok" if ok else "FAILED")


@pytest.mark.parametrize("consume_every", [1, 4])
def test_p2p_gather_with_a_slow_rank(tmp_path, oracle, consume_every):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    result = str(tmp_path / "r")
    port = 29600 + consume_every
    mp.spawn(_worker, args=(port, consume_every, result), nprocs=2, join=True)
    for r in range(2):
        assert open(result + ".%d" % r).read() == "ok"
