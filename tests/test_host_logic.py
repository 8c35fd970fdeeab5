"""CPU: host-side logic added around the kernels -- pair ordering of the batched inference driver, the bit-reproducible
synthetic sequences the full-size fixtures are regenerated from, the fixtures' own consistency."""
import hashlib
import os

import numpy as np
import torch

from conftest import GOLDEN
from rvdd_release_b200 import synth


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_sequence_pairs_follow_the_dataset_order():
    """data/infer4rec_dataset.py:198-202: per frame t the flows [t-1 -> t, t+1 -> t]; the driver batches all past pairs of all
    sequences, then all future pairs, indices into the flattened [S * T] frame list."""
    from rvdd_release_b200 import infer
    s, t = infer.sequence_pairs(2, 5, 1)
    assert s.tolist() == [0, 1, 2, 5, 6, 7, 2, 3, 4, 7, 8, 9] and t.tolist() == [1, 2, 3, 6, 7, 8] * 2
    s, t = infer.sequence_pairs(1, 4, 0)
    assert s.tolist() == [0, 1, 2] and t.tolist() == [1, 2, 3]
    s, t = infer.sequence_pairs(3, 2, 0)
    assert s.tolist() == [0, 2, 4] and t.tolist() == [1, 3, 5]


def test_exact_sequence_matches_the_fixture_hash():
    """synth.exact_sequence uses only int64 hashing and IEEE + - * / floor sqrt: the frames hash to what the build container
    produced when the fixture was made (the GPU test asserts the same on the device)."""
    d = np.load(os.path.join(GOLDEN, "config_cn_small.npz"))
    nfr, h, w = (int(v) for v in d["geometry"])
    seq = synth.exact_sequence(nfr, h, w, str(d["iso"]))
    assert seq.dtype == torch.float32 and tuple(seq.shape) == (nfr, h, w, 4)
    assert _sha(seq.numpy()) == str(d["sha_frames"])
    again = synth.exact_sequence(nfr, h, w, str(d["iso"]))
    assert torch.equal(seq, again)
    other = synth.exact_sequence(nfr, h, w, str(d["iso"]), noise_seed=1)
    assert not torch.equal(seq, other) and float((seq - other).abs().mean()) > 1.0        # another noise realisation
    clean = synth.exact_sequence(2, h, w, "clean")
    assert float(clean.min()) >= synth.ISO["clean"]["lo"] and float(clean.max()) <= synth.ISO["clean"]["hi"]


def test_exact_gray_pair_matches_the_large_golden_recipe():
    """The 3840x2160 golden's inputs come from synth.exact_gray_pair; a small instance must be deterministic and textured."""
    a0, a1 = synth.exact_gray_pair(45, 80)
    b0, b1 = synth.exact_gray_pair(45, 80)
    assert np.array_equal(a0, b0) and np.array_equal(a1, b1) and a0.dtype == np.float32
    assert float(a0.std()) > 50.0 and not np.array_equal(a0, a1)


def test_config_fixtures_are_self_consistent():
    """Every full-size fixture holds one PSNR, one mean and one flow hash per denoised frame (and per future flow)."""
    for name, nden in (("c2", 29), ("c3", 28), ("c5", 3), ("cn_small", 4)):
        d = np.load(os.path.join(GOLDEN, "config_%s.npz" % name))
        assert d["psnr"].shape == (nden,) and d["denoised_mean"].shape == (nden,) and d["flow_sha"].shape == (nden,)
        assert d["flow_sub"].shape[0] == nden and np.isfinite(d["psnr"]).all()
        if name != "c2":
            assert d["future_flow_sha"].shape == (nden,)
        assert len(set(d["flow_sha"].tolist())) == nden                 # every pair has its own flow
