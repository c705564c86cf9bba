"""SURVEY.md section 8f row 1 (second half): 2-D label maps from rendered query-class logits (siu3r_b200.labels2d <-> src/pipeline.py:132-193,
viewer.py:422-435) -- the CUDA kernel through the C-ABI against
  * tests/golden/labels2d_cases.npz: outputs of the reference's own statements (oracle/make_golden_labels2d.py), bit-exact, and
  * the oracle restatement (oracle/labels2d_ref.py, pinned to those goldens in tests/test_oracle_cpu.py) on larger seeded inputs and on
    logits that come out of SplattingCUDA.forward(render_qc_logits=True).
Integer outputs: everything is compared for equality.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(GOLD, "labels2d_cases.npz"), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("layout", ["channel_last_view", "contiguous"])
def test_labels2d_matches_reference_goldens(layout):
    from siu3r_b200.labels2d import labels_from_qc_logits
    z, meta = _cases()
    for name, m in meta.items():
        x = torch.from_numpy(z[name + "__logits"]).to(DEV)
        if layout == "channel_last_view":                 # what the rasteriser hands over: [v, h, w, q, c] memory viewed as [v, q, c, h, w]
            x = x.permute(0, 3, 4, 1, 2).contiguous().permute(0, 3, 4, 1, 2)
            assert not x.is_contiguous()
        sem, ins, infos = labels_from_qc_logits([x], [m["scores"]], m["label_ids_to_fuse"], m["num_queries"])
        assert sem.dtype == torch.int64 and ins.dtype == torch.int64 and tuple(sem.shape) == (1, *z[name + "__sem"].shape)
        assert np.array_equal(sem[0].cpu().numpy(), z[name + "__sem"]), name
        assert np.array_equal(ins[0].cpu().numpy(), z[name + "__ins"]), name
        assert infos[0] == m["infos"], (name, infos[0], m["infos"])


@pytest.mark.parametrize("v,q,c,h,w", [(2, 12, 21, 96, 128), (1, 30, 21, 256, 256), (3, 2, 5, 33, 17), (1, 1, 70, 20, 20)])
def test_labels2d_vs_oracle_seeded(v, q, c, h, w):
    """Larger seeded inputs, batch of two samples with different query counts, against the oracle; also the viewer variant and a second
    call (idempotence: the kernel leaves no state behind)."""
    from oracle import labels2d_ref as LR
    from siu3r_b200.labels2d import labels_from_qc_logits, viewer_labels
    g = torch.Generator().manual_seed(v * 1000 + q)
    xs, scores = [], []
    for qq in (q, max(1, q // 2)):
        x = torch.rand(v, h, w, qq, c, generator=g) * 0.28
        for qi in range(qq):                              # one confident disc per query and view
            cy, cx, r = torch.rand(3, generator=g)
            yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            disc = ((yy - cy * h) ** 2 + (xx - cx * w) ** 2) < (0.1 + 0.3 * r) ** 2 * h * w
            x[:, :, :, qi, (qi * 5 + 1) % c] += disc[None].float() * (0.2 + 0.7 * float(torch.rand(1, generator=g)))
        x = torch.round(x * 64) / 64                      # coarse values -> plenty of exact ties
        xs.append(x.to(DEV).permute(0, 3, 4, 1, 2))       # [v, q, c, h, w] view of channel-last memory
        scores.append([0.55 + 0.01 * i for i in range(qq)])
    sem, ins, infos = labels_from_qc_logits(xs, scores, (0, 1), 100)
    sem2, ins2, infos2 = labels_from_qc_logits(xs, scores, (0, 1), 100)
    assert torch.equal(sem, sem2) and torch.equal(ins, ins2) and infos == infos2
    for bi, (x, sc) in enumerate(zip(xs, scores)):
        rs, ri, rinfo = LR.labels_from_qc_logits(x.cpu().numpy(), sc, (0, 1), 100)
        assert np.array_equal(sem[bi].cpu().numpy(), rs) and np.array_equal(ins[bi].cpu().numpy(), ri)
        assert infos[bi] == rinfo
    vs, vi = viewer_labels(xs[0], 0.3)
    rs, ri, _ = LR.labels_from_qc_logits(xs[0].cpu().numpy(), scores[0], (), 100)
    ri = ri.copy()
    for i in (1, 2):                                      # viewer.py:433-434
        ri[rs == i] = 100 + i + 1
    assert np.array_equal(vs.cpu().numpy(), rs) and np.array_equal(vi.cpu().numpy(), ri)


def test_labels2d_after_splatting_and_error_behaviour():
    """The kernel consumes what SplattingCUDA.forward(render_qc_logits=True) returns (a permuted view of the rasteriser's channel-last
    buffer) without a copy; bad arguments are refused by the C-ABI (no exception-free silent path)."""
    from oracle import labels2d_ref as LR
    from siu3r_b200 import _lib
    from siu3r_b200.gaussians import Gaussians
    from siu3r_b200.labels2d import labels_from_qc_logits
    from siu3r_b200.renderer import SplattingCUDA
    G, Q, C, S = 6000, 3, 21, 64
    rng = np.random.default_rng(11)
    z = rng.uniform(0.3, 2.0, G)
    means = np.stack([rng.uniform(-1, 1, G) * z * 0.4, rng.uniform(-1, 1, G) * z * 0.4, z], -1).astype("f4")
    A = (rng.standard_normal((G, 3, 3)) * (0.02 * z)[:, None, None]).astype("f4")
    cov = (A @ A.transpose(0, 2, 1) + 1e-7 * np.eye(3, dtype="f4")).astype("f4")
    t = lambda a: torch.from_numpy(a).to(DEV)
    g = Gaussians(means=t(means)[None], covariances=t(cov)[None], harmonics=torch.zeros(1, G, 3, 25, device=DEV), opacities=torch.full((1, G), 0.9, device=DEV))
    probs = torch.rand(G, Q, C, generator=torch.Generator().manual_seed(1)).to(DEV)
    g.seg_query_class_logits = [probs]
    E = torch.eye(4)[None, None].repeat(1, 2, 1, 1)
    E[0, 1, 0, 3] = 0.01
    K = torch.tensor([[1.242, 0, 0.5], [0, 1.242, 0.5], [0, 0, 1.0]])[None, None].repeat(1, 2, 1, 1)
    out = SplattingCUDA()(g, E, K, (S, S), render_color=False, render_qc_logits=True)
    qc = out["render_qc_logits"]
    assert tuple(qc[0].shape) == (2, Q, C, S, S) and not qc[0].is_contiguous()
    scores = [[0.9, 0.8, 0.7]]
    sem, ins, infos = labels_from_qc_logits(qc, scores, (0, 1), 100)
    rs, ri, rinfo = LR.labels_from_qc_logits(qc[0].cpu().numpy(), scores[0], (0, 1), 100)
    assert np.array_equal(sem[0].cpu().numpy(), rs) and np.array_equal(ins[0].cpu().numpy(), ri) and infos[0] == rinfo
    assert int((sem > 0).sum()) > 0                       # the scene really produces labelled pixels
    lib = _lib.load()
    assert lib.siu3r_labels_from_qc_logits(None, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0.3, None, None, 0, None, None, None, None, None) != 0
    with pytest.raises(AssertionError):
        labels_from_qc_logits([qc[0].cpu()], scores)      # host tensors are refused: there is no CPU path
