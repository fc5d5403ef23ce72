"""Randomised shapes / parameters for MultiBoxDetection and MultiBoxTarget against the CPU oracle (bit-exact), aimed at
the size-dependent paths the preset tests do not reach: anchor counts that are not multiples of 4, class counts outside
the specialised kernels, more than 1024 tiles per image, keys that do not fit in shared memory, nms_topk <= 0, many
label slots.  Run under gpurun:  python scripts/fuzz_parity.py [cases]"""
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from dspnet_b200 import MultiBoxDetection, MultiBoxTarget
from oracle import oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
huge = len(sys.argv) > 2 and sys.argv[2] == 'huge'  # only anchor counts beyond 1024 tiles / the shared-memory key budget
dev = torch.device('cuda', 0)
r = np.random.Generator(np.random.PCG64(2026))


def anchors_of(A):
    c = r.uniform(0.05, 0.95, (A, 2))
    wh = r.uniform(0.02, 0.5, (A, 2))
    return np.concatenate([c - wh / 2, c + wh / 2], axis=1).astype(np.float32)[None]


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


bad = 0
for case in range(n_cases):
    big = huge or case % 6 == 5
    A = int(r.choice([262148 + 4 * int(r.integers(0, 3)), 300000, 270001] if huge else [37, 1000, 4097, 30001])) if big else int(r.integers(5, 6000))
    C = int(r.choice([3, 21] if huge else [2, 3, 9, 21, 30]))
    B = int(r.integers(1, 3 if huge else 4))
    an = anchors_of(A)
    # detection: sparse foreground so that the oracle's O(V^2) NMS stays cheap for big A
    logits = r.standard_normal((B, C, A)).astype(np.float32)
    logits[:, 0] += np.where(r.random((B, A)) < (0.995 if A > 20000 else 0.8), 8.0, 0.0).astype(np.float32)
    e = np.exp(logits - logits.max(1, keepdims=True))
    prob = (e / e.sum(1, keepdims=True)).astype(np.float32)
    loc = np.concatenate([r.normal(0, 0.5, (B, A, 4)), r.uniform(0, 10, (B, A, 1))], axis=2).astype(np.float32).reshape(B, A * 5)
    kw = dict(threshold=float(r.choice([0.01, 0.2])), clip=bool(r.integers(0, 2)), nms_threshold=float(r.choice([0.3, 0.45, 0.7, -1.0])),
              force_suppress=bool(r.integers(0, 2)), nms_topk=int(r.choice([-1, 1, 50, 400, 1000])))
    want = O.multibox_detection(prob, loc, an, **kw)
    got = MultiBoxDetection(t(prob), t(loc), t(an), **kw).cpu().numpy()
    ok_d = np.array_equal(want.view(np.uint32), got.view(np.uint32))
    # target
    L = int(r.choice([1, 8, 58, 200]))
    lab = np.full((B, L, 6), -1, np.float32)
    for b in range(B):
        g = int(r.integers(0, L + 1))
        c = r.uniform(0.1, 0.9, (g, 2))
        wh = r.uniform(0.03, 0.6, (g, 2))
        lab[b, :g, 0] = r.integers(0, C - 1, g) if C > 1 else 0
        lab[b, :g, 1:3] = np.clip(c - wh / 2, 0, 1)
        lab[b, :g, 3:5] = np.clip(c + wh / 2, 0, 1)
        lab[b, :g, 5] = r.uniform(0, 1, g)
    tkw = dict(overlap_threshold=float(r.choice([0.5, 0.3])), negative_mining_ratio=float(r.choice([-1.0, 3.0])),
               negative_mining_thresh=0.5)
    try:
        twant = O.multibox_target(an, lab, logits, **tkw)
        tgot = [x.cpu().numpy() for x in MultiBoxTarget(t(an), t(lab), t(logits), **tkw)]
        ok_t = all(np.array_equal(np.asarray(w, np.float32).view(np.uint32), g.reshape(np.asarray(w).shape).view(np.uint32))
                   for w, g in zip(twant, tgot))
    except Exception as ex:  # both sides must raise the same data-dependent error
        try:
            MultiBoxTarget(t(an), t(lab), t(logits), **tkw)
            ok_t = False
        except Exception as ex2:
            ok_t = type(ex).__name__[-5:] == type(ex2).__name__[-5:]
    bad += (not ok_d) + (not ok_t)
    print("case %2d A=%6d C=%2d B=%d L=%3d det %s (%s) tgt %s" % (case, A, C, B, L, "ok" if ok_d else "MISMATCH", kw, "ok" if ok_t else "MISMATCH"), flush=True)
print("FUZZ mismatches:", bad)
