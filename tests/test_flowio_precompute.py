"""CPU: on-disk flow format (TIFF without libtiff) and the sharded precompute driver (host logic only; the compute
function is injected, so no GPU is needed).  The multi-rank path runs with world_size 2 over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from rvdd_release_b200 import flowio, precompute


def test_tif_round_trip_all_channel_counts(tmp_path):
    rng = np.random.RandomState(0)
    for c in (1, 2, 3, 4):
        a = rng.randn(37, 53, c).astype(np.float32)
        p = str(tmp_path / ("x%d.tif" % c))
        flowio.write_tif(p, a)
        b = flowio.read_tif(p)
        assert b.dtype == np.float32 and b.shape == a.shape and np.array_equal(a, b)
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]          # atomic write leaves nothing behind


def test_tif_is_readable_by_an_independent_tiff_reader(tmp_path):
    """libtiff is what the reference's loaders use (through iio); OpenCV bundles libtiff, so let it read our files."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(1)
    a1 = rng.randn(21, 34).astype(np.float32)
    flowio.write_tif(str(tmp_path / "g.tif"), a1)
    assert np.array_equal(cv2.imread(str(tmp_path / "g.tif"), cv2.IMREAD_UNCHANGED), a1)
    a4 = rng.randn(21, 34, 4).astype(np.float32)
    flowio.write_tif(str(tmp_path / "q.tif"), a4)
    b4 = cv2.imread(str(tmp_path / "q.tif"), cv2.IMREAD_UNCHANGED)
    assert b4 is not None and np.array_equal(b4[:, :, [2, 1, 0, 3]], a4)     # OpenCV hands back BGRA


def test_reads_lzw_tiffs_like_the_reference_writes(tmp_path):
    """iio.write LZW-compresses images below 2000x2000 (iio.c:3001-3004); OpenCV's writer does the same by default."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(2)
    smooth = np.cumsum(rng.randn(64, 96).astype(np.float32), axis=1)
    for name, img in (("noise", rng.randn(40, 50).astype(np.float32)), ("smooth", smooth),
                      ("const", np.full((33, 47), 1.5, np.float32))):
        p = str(tmp_path / (name + ".tif"))
        assert cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, 5])
        assert np.array_equal(flowio.read_tif(p)[:, :, 0], img), name


def test_video_pairs_follow_the_reference_loops():
    # patch_depth 2: flow from frame z to z+1; future depth 1: flow from z+1 to z
    assert precompute.video_pairs(4, 2, 0) == [(0, 1), (1, 2), (2, 3)]
    assert precompute.video_pairs(4, 2, 1) == [(0, 1), (1, 2), (2, 3), (1, 0), (2, 1), (3, 2)]
    assert precompute.video_pairs(4, 3, 0) == [(0, 2), (1, 2), (1, 3), (2, 3)]
    assert precompute.video_pairs(1, 2, 1) == []


def _fake_compute(frames, src, tgt, want_warp):
    """Stands in for the GPU: flow = (mean of target - mean of source, source index) so results are checkable."""
    k, (_, h, w, c) = len(src), frames.shape
    flow = np.zeros((k, h, w, 2), np.float32)
    for i, (s, t) in enumerate(zip(src, tgt)):
        flow[i, :, :, 0] = frames[t].mean() - frames[s].mean()
        flow[i, :, :, 1] = frames[s, 0, 0, 0]
    return flow, (frames[list(src)].copy() if want_warp else None)


def _make_dataset(root, nvid=5, nfr=4):
    videos = []
    for v in range(nvid):
        d = os.path.join(root, "noisy", "vid%02d" % v)
        os.makedirs(d)
        paths = []
        for f in range(nfr):
            p = os.path.join(d, "%04d.tif" % f)
            flowio.write_tif(p, np.full((6, 8, 4), 100 * v + f, np.float32))
            paths.append(p)
        videos.append(("vid%02d" % v, paths))
    return videos


def test_precompute_writes_reference_layout_and_resumes(tmp_path):
    videos = _make_dataset(str(tmp_path))
    flow_root = str(tmp_path / "flow")
    files = precompute.precompute_dataset(videos, flow_root, None, 2, 1, compute=_fake_compute)
    assert len(files) == 5 * 6
    f = flowio.read_tif(os.path.join(flow_root, "vid03", "0001_0002.tif"))     # <fromCode>_<toCode>.tif (library.py:140)
    assert f.shape == (6, 8, 2) and f[0, 0, 0] == 1.0 and f[0, 0, 1] == 301.0
    g = flowio.read_tif(os.path.join(flow_root, "vid03", "0002_0001.tif"))     # future pair
    assert g[0, 0, 0] == -1.0 and g[0, 0, 1] == 302.0
    assert precompute.list_videos(str(tmp_path / "noisy")) == videos
    # resume: nothing left to do; delete one file and only that one is recomputed
    assert precompute.precompute_dataset(videos, flow_root, None, 2, 1, compute=_fake_compute) == []
    os.remove(os.path.join(flow_root, "vid01", "0000_0001.tif"))
    assert precompute.precompute_dataset(videos, flow_root, None, 2, 1, compute=_fake_compute) == \
        [os.path.join(flow_root, "vid01", "0000_0001.tif")]


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from rvdd_release_b200 import precompute
from test_flowio_precompute import _fake_compute
root = sys.argv[2]
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[3], rank=int(sys.argv[4]), world_size=2)
videos = precompute.list_videos(os.path.join(root, "noisy"))
files = precompute.precompute_dataset(videos, os.path.join(root, "flow"), None, 2, 1, rank=dist.get_rank(), world=2,
                                      compute=_fake_compute, gather=precompute.torch_gather)
mine = [f for f in files if ("vid%02d" % 0 in f or "vid02" in f or "vid04" in f)] if dist.get_rank() == 0 else None
print("RANK", dist.get_rank(), len(files), len(set(files)))
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    videos = _make_dataset(str(tmp_path))
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, "-c", _WORKER, ROOT, str(tmp_path), port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    for o, _ in outs:
        assert o.strip().split()[2:] == ["30", "30"]            # every rank sees all 30 files after the gather
    assert precompute.shard(list(range(5)), 0, 2) == [0, 2, 4] and precompute.shard(list(range(5)), 1, 2) == [1, 3]
    flow_root = str(tmp_path / "flow")
    assert sum(len(os.listdir(os.path.join(flow_root, v))) for v, _ in videos) == 30


def test_read_tif_into_fast_path_and_fallback(tmp_path):
    rng = np.random.RandomState(3)
    a = rng.randn(19, 23, 4).astype(np.float32)
    p = str(tmp_path / "a.tif")
    flowio.write_tif(p, a)
    out = np.zeros((19, 23, 4), np.float32)
    assert flowio.read_tif_into(p, out) and np.array_equal(out, a)
    assert not flowio.read_tif_into(p, np.zeros((19, 23, 3), np.float32))            # wrong shape -> caller falls back
    raw = bytearray(open(p, "rb").read())
    ifd = int.from_bytes(raw[4:8], "little")
    for i in range(int.from_bytes(raw[ifd:ifd + 2], "little")):
        e = ifd + 2 + 12 * i
        if int.from_bytes(raw[e:e + 2], "little") == 259:                              # Compression := LZW
            raw[e + 8:e + 10] = (5).to_bytes(2, "little")
    q = str(tmp_path / "marked_lzw.tif")
    open(q, "wb").write(raw)
    assert not flowio.read_tif_into(q, np.zeros((19, 23, 4), np.float32))            # compressed -> caller falls back
    cv2 = pytest.importorskip("cv2")
    strips = str(tmp_path / "strips.tif")
    assert cv2.imwrite(strips, a[:, :, 0], [cv2.IMWRITE_TIFF_COMPRESSION, 1, cv2.IMWRITE_TIFF_ROWSPERSTRIP, 4])
    out1 = np.zeros((19, 23, 1), np.float32)
    assert flowio.read_tif_into(strips, out1) and np.array_equal(out1[:, :, 0], a[:, :, 0])   # libtiff's multi-strip layout


def test_plan_orders_past_and_future_pairs_by_frame(tmp_path):
    videos = _make_dataset(str(tmp_path), nvid=1, nfr=5)
    todo = precompute.plan_video(videos[0][1], str(tmp_path / "flow" / "vid00"), None, 2, 1)
    assert [(p["src"], p["tgt"]) for p in todo] == [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2), (3, 4), (4, 3)]
    assert all(p["need_flow"] for p in todo)


def test_cached_flow_is_reused_for_a_missing_warp(tmp_path):
    """base_dataset.py:182-185: flow file present, warped file missing -> warp with the CACHED flow, no recomputation."""
    videos = _make_dataset(str(tmp_path), nvid=1, nfr=3)
    flow_root, warp_root = str(tmp_path / "flow"), str(tmp_path / "warped")
    precompute.precompute_dataset(videos, flow_root, None, 2, 0, compute=_fake_compute)
    marker = np.full((6, 8, 2), 7.0, np.float32)
    flowio.write_tif(os.path.join(flow_root, "vid00", "0000_0001.tif"), marker)       # a cache written by someone else
    calls = []

    def fake_warp(img, flow):
        calls.append(float(flow[0, 0, 0]))
        return img + flow[:, :, :1]

    def no_compute(*a):
        raise AssertionError("the flow must not be recomputed")

    os.makedirs(os.path.join(warp_root, "vid00"))
    files = precompute.precompute_video(videos[0][1], os.path.join(flow_root, "vid00"), os.path.join(warp_root, "vid00"),
                                        2, 0, compute=no_compute, warp=fake_warp)
    assert len(files) == 2 and calls[0] == 7.0
    w = flowio.read_tif(os.path.join(warp_root, "vid00", "0000_0001.tif"))
    assert w[0, 0, 0] == 0.0 + 7.0
    assert np.array_equal(flowio.read_tif(os.path.join(flow_root, "vid00", "0000_0001.tif")), marker)


class _FakeBridge:
    """Stands in for libBridge.so's two staging slots: `submit_host` computes on a worker thread, `wait_host` joins it --
    so the pipeline's overlap logic (reads ahead, deferred writes, buffer reuse) runs without a GPU."""

    def __init__(self):
        import threading
        self.threading, self.jobs, self.order = threading, [None, None], []

    def submit_host(self, slot, frames, src, tgt, flow_out, warped_out=None, params=None):
        assert self.jobs[slot] is None, "slot submitted twice without a wait"

        def work():
            f, w = _fake_compute(frames.numpy(), src, tgt, warped_out is not None)
            flow_out.numpy()[...] = f
            if warped_out is not None:
                warped_out.numpy()[...] = w

        t = self.threading.Thread(target=work)
        t.start()
        self.jobs[slot] = t
        self.order.append(("submit", slot))

    def wait_host(self, slot):
        if self.jobs[slot] is not None:
            self.jobs[slot].join()
            self.jobs[slot] = None
        self.order.append(("wait", slot))


def test_pipelined_driver_with_reader_and_writer_pools(tmp_path):
    """The GPU path's pipeline (reader pool -> staging slots -> writer pool) against the synchronous driver: same files,
    same contents, for videos of different lengths (ragged tail batches) and a batch size that splits videos."""
    root = str(tmp_path)
    videos = []
    for v, nfr in enumerate((7, 3, 12, 2, 5)):
        d = os.path.join(root, "noisy", "vid%02d" % v)
        os.makedirs(d)
        paths = []
        for f in range(nfr):
            p = os.path.join(d, "%04d.tif" % f)
            flowio.write_tif(p, np.full((6, 8, 4), 100 * v + f, np.float32))
            paths.append(p)
        videos.append(("vid%02d" % v, paths))
    ref_root, pipe_root = os.path.join(root, "flow_ref"), os.path.join(root, "flow_pipe")
    ref_files = precompute.precompute_dataset(videos, ref_root, os.path.join(root, "w_ref"), 2, 1, compute=_fake_compute,
                                              max_pairs_per_batch=5)
    stats = {}
    fb = _FakeBridge()
    files = precompute.precompute_dataset(videos, pipe_root, os.path.join(root, "w_pipe"), 2, 1, max_pairs_per_batch=5,
                                          readers=3, writers=3, stats=stats, bridge=fb)
    rel = lambda fs, base: sorted(os.path.relpath(f, root).replace(base, "X") for f in fs)
    assert rel(files, "_pipe") == rel(ref_files, "_ref") and len(files) == 2 * 2 * (6 + 2 + 11 + 1 + 4)
    for f in ref_files:
        g = f.replace("_ref", "_pipe")
        assert np.array_equal(flowio.read_tif(f), flowio.read_tif(g)), g
    assert stats["pairs"] == 48 and stats["batches"] >= 10 and stats["pinned_bytes"] > 0
    # pinned footprint is bounded by NHOST x the largest batch (6 frames + 5 flows + 5 warped frames of 6x8 pixels)
    assert stats["pinned_bytes"] <= 4 * 4 * (6 * 6 * 8 * 4 + 5 * 6 * 8 * 2 + 5 * 6 * 8 * 4)
    # a slot is always waited for before it is submitted again
    busy = [False, False]
    for what, slot in fb.order:
        if what == "submit":
            assert not busy[slot]
            busy[slot] = True
        else:
            busy[slot] = False
    # resume: nothing left
    assert precompute.precompute_dataset(videos, pipe_root, os.path.join(root, "w_pipe"), 2, 1, bridge=fb) == []
