"""Offline flow precompute: the flow cache the reference builds in its dataset constructors, on GPUs.

Mirrors ``createWarpedInputData`` / ``createFutureWarpedInputData`` (data/base_dataset.py:134-249): for every video and
every target frame, the flows from the ``patch_depth-1`` previous frames and from the ``future_patch_depth`` next
frames to the target are computed (unless their file already exists) and written as
``<flow_root>/<video>/<fromCode>_<toCode>.tif`` with the ``(h, w, 2)`` float32 content ``iio.write`` would produce
(library.py:140-141, base_dataset.py:152-180).  ``gen_warp`` additionally stores the warped source frames.

Where the reference runs one pair at a time on the CPU, this driver hands whole videos to the persistent GPU solver
(all pairs of a video in one batch) and shards the videos over ranks: one process per GPU, rank r takes videos
r, r+world, ... -- the pairs are independent, so there is no collective on the data path; the ranks only gather the
list of files they wrote at the end.
"""
import os
import time

import numpy as np

from . import flowio
from .library import warpedimagefile


def video_pairs(n_frames, patch_depth=2, future_patch_depth=0):
    """(source, target) frame-index pairs of one video, in the reference's order.

    Past pairs (base_dataset.py:147-165): target z+PD-1, sources z .. z+PD-2, for z in range(n-PD+1).
    Future pairs (base_dataset.py:205-223): target z, sources z+1 .. z+FD, for z in range(n-FD)."""
    pairs = []
    for z in range(n_frames - patch_depth + 1):
        for n in range(patch_depth - 1):
            pairs.append((z + n, z + patch_depth - 1))
    if future_patch_depth > 0:
        for z in range(n_frames - future_patch_depth):
            for n in range(future_patch_depth):
                pairs.append((z + n + 1, z))
    return pairs


def shard(items, rank, world):
    """Round-robin shard of a list of videos over the ranks."""
    return [it for i, it in enumerate(items) if i % world == rank]


def _code(path):
    return os.path.splitext(os.path.basename(path))[0]


def plan_video(frame_paths, flow_dir, warp_dir=None, patch_depth=2, future_patch_depth=0):
    """The pairs of one video that still need work -> list of dicts(src, tgt, flow_file, warp_file, need_flow), ordered
    by frame position so that a batch's past and future pairs share their frames (one upload for both)."""
    todo = []
    for s, t in video_pairs(len(frame_paths), patch_depth, future_patch_depth):
        ff = warpedimagefile(flow_dir, _code(frame_paths[s]), _code(frame_paths[t]))
        wf = warpedimagefile(warp_dir, _code(frame_paths[s]), _code(frame_paths[t])) if warp_dir else None
        need_flow = not os.path.isfile(ff)
        if need_flow or (wf and not os.path.isfile(wf)):
            todo.append(dict(src=s, tgt=t, flow_file=ff, warp_file=wf, need_flow=need_flow))
    todo.sort(key=lambda p: (min(p["src"], p["tgt"]), p["src"] > p["tgt"]))
    return todo


def gpu_compute(frames, src, tgt, want_warp):
    """Default compute function: one batched call into libBridge.so with host buffers."""
    from . import bridge
    flow, warped, _ = bridge.default_bridge().flow_and_warp_host(frames, src, tgt, want_warp=want_warp)
    return flow.numpy(), (warped.numpy() if want_warp else None)


def warp_only(frame, flow):
    """A pair whose flow file exists but whose warped file is missing: warp with the CACHED flow, exactly as
    base_dataset.py:182-185 does (the cache may have been written by the reference or with other parameters)."""
    from . import flow_utils
    return flow_utils.single_warp(frame, flow)


def _split_cached(todo, frame_paths, written, warp=warp_only):
    """Handle the warp-only pairs of a plan on the spot and return the pairs that need a flow."""
    rest = []
    for p in todo:
        if p["need_flow"]:
            rest.append(p)
            continue
        flow = flowio.read_tif(p["flow_file"]).astype(np.float32)
        img1 = flowio.read_image(frame_paths[p["src"]]).astype(np.float32)
        flowio.write_tif(p["warp_file"], warp(img1, flow))                                    # base_dataset.py:189
        written.append(p["warp_file"])
    return rest


def precompute_video(frame_paths, flow_dir, warp_dir=None, patch_depth=2, future_patch_depth=0, compute=gpu_compute,
                     max_pairs_per_batch=64, warp=warp_only):
    """Create the missing flow (and optionally warped) files of one video.  Returns the files written."""
    todo = plan_video(frame_paths, flow_dir, warp_dir, patch_depth, future_patch_depth)
    if not todo:
        return []
    os.makedirs(flow_dir, exist_ok=True)
    if warp_dir:
        os.makedirs(warp_dir, exist_ok=True)
    written = []
    todo = _split_cached(todo, frame_paths, written, warp)
    for b0 in range(0, len(todo), max_pairs_per_batch):
        batch = todo[b0:b0 + max_pairs_per_batch]
        used = sorted({p["src"] for p in batch} | {p["tgt"] for p in batch})
        local = {f: i for i, f in enumerate(used)}
        frames = np.stack([flowio.read_image(frame_paths[f]).astype(np.float32) for f in used])   # base_dataset.py:159,174
        src = [local[p["src"]] for p in batch]
        tgt = [local[p["tgt"]] for p in batch]
        flow, warped = compute(frames, src, tgt, warp_dir is not None)
        for k, p in enumerate(batch):
            flowio.write_tif(p["flow_file"], flow[k])                                         # base_dataset.py:180
            written.append(p["flow_file"])
            if p["warp_file"] and not os.path.isfile(p["warp_file"]):
                flowio.write_tif(p["warp_file"], warped[k])                                   # base_dataset.py:189
                written.append(p["warp_file"])
    return written


class _HostSet:
    """One set of pinned host buffers (frames in, flows out, warped frames out).  There are NHOST of them per process,
    each sized for the largest batch seen so far; a bigger batch frees the old buffers before allocating new ones, so
    the pinned footprint is bounded by NHOST x the largest batch, whatever the mix of video lengths."""

    def __init__(self, torch):
        self.t = torch
        self.pin = torch.cuda.is_available()               # (host-logic tests drive the pipeline with a stand-in bridge)
        self.frames = self.flow = self.warp = None
        self.writes = []                                   # futures of the files being written from this set

    def _fit(self, name, shape):
        n = int(np.prod(shape))
        buf = getattr(self, name)
        if buf is None or buf.numel() < n:
            setattr(self, name, None)                      # release the old pinned block first
            del buf
            buf = self.t.empty(n, dtype=self.t.float32)
            setattr(self, name, buf.pin_memory() if self.pin else buf)

    def shape(self, nfr, h, w, c, npairs, want_warp):
        self._fit("frames", (nfr, h, w, c))
        self._fit("flow", (npairs, h, w, 2))
        if want_warp:
            self._fit("warp", (npairs, h, w, c))
        return (self.frames[:nfr * h * w * c].view(nfr, h, w, c), self.flow[:npairs * h * w * 2].view(npairs, h, w, 2),
                self.warp[:npairs * h * w * c].view(npairs, h, w, c) if want_warp else None)

    def pinned_bytes(self):
        return 4 * sum(b.numel() for b in (self.frames, self.flow, self.warp) if b is not None)


class _PipelinedGPU:
    """The offline precompute as a pipeline: a pool of reader threads decodes the frame files of
    batch i+1 straight into pinned memory while the GPU (libBridge.so's two staging slots,
    rvdd_flow_and_warp_host_submit / _wait: upload, kernels and download of consecutive batches overlap on three streams)
    works on batches i and i-1, and a pool of writer threads stores the flows of batch i-2.  The submitting thread only
    schedules; file I/O never sits between two GPU submissions."""
    NHOST = 4

    def __init__(self, readers=None, writers=None, bridge=None):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        if bridge is None:
            from . import bridge as _b
            bridge = _b.default_bridge()
        self.torch, self.br = torch, bridge
        # default pool sizes: the cores this process may use, shared with the other ranks of the node, half for each pool
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
        ncpu = max(1, ncpu // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        self.readers = ThreadPoolExecutor(readers or max(2, min(8, ncpu // 2)), thread_name_prefix="rvdd-read")
        self.writers = ThreadPoolExecutor(writers or max(2, min(8, ncpu // 2)), thread_name_prefix="rvdd-write")
        self.sets = [_HostSet(torch) for _ in range(self.NHOST)]
        self.k = 0
        self.reading = None                                # (job, read futures) of the batch being read
        self.on_gpu = [None, None]                         # job per staging slot
        self.nsub = 0
        self.written = []
        self.stats = dict(batches=0, pairs=0, frames_read=0, bytes_read=0, bytes_written=0, wait_read_s=0.0, wait_gpu_s=0.0,
                          wait_write_s=0.0)         # where the submitting thread waited: readers, the GPU, writers

    # ---- stage 1: read
    def _start_read(self, frame_paths, batch, want_warp):
        hs = self.sets[self.k % self.NHOST]
        self.k += 1
        t0 = time.perf_counter()
        for f in hs.writes:                                # the files of the batch that used this set 4 batches ago
            f.result()
        self.stats["wait_write_s"] += time.perf_counter() - t0
        hs.writes = []
        used = sorted({p["src"] for p in batch} | {p["tgt"] for p in batch})
        local = {f: i for i, f in enumerate(used)}
        first = flowio.read_image(frame_paths[used[0]])
        h, w, c = first.shape
        frames, flow, warped = hs.shape(len(used), h, w, c, len(batch), want_warp)
        fr_np = frames.numpy()
        fr_np[0] = first

        def load(i, path):                                 # pixels straight into pinned memory (one readinto) ...
            if not flowio.read_tif_into(path, fr_np[i]):
                fr_np[i] = flowio.read_image(path)         # ... or decode + float32 conversion for other formats

        futs = [self.readers.submit(load, i, frame_paths[f]) for i, f in enumerate(used[1:], 1)]
        self.stats["frames_read"] += len(used)
        self.stats["bytes_read"] += frames.numel() * 4
        job = dict(hs=hs, batch=batch, frames=frames, flow=flow, warped=warped,
                   src=[local[p["src"]] for p in batch], tgt=[local[p["tgt"]] for p in batch])
        return job, futs

    # ---- stage 2: GPU
    def _submit(self, job, futs):
        t0 = time.perf_counter()
        for f in futs:
            f.result()                                     # frames of this batch are in pinned memory
        self.stats["wait_read_s"] += time.perf_counter() - t0
        slot = self.nsub & 1
        self.nsub += 1
        self._retire(slot)                                 # the batch submitted two steps ago on this slot
        self.br.submit_host(slot, job["frames"], job["src"], job["tgt"], job["flow"], job["warped"])
        self.on_gpu[slot] = job

    # ---- stage 3: write
    def _retire(self, slot):
        job = self.on_gpu[slot]
        if job is None:
            return
        t0 = time.perf_counter()
        self.br.wait_host(slot)
        self.stats["wait_gpu_s"] += time.perf_counter() - t0
        self.on_gpu[slot] = None
        flow, warped = job["flow"].numpy(), (job["warped"].numpy() if job["warped"] is not None else None)
        for k, p in enumerate(job["batch"]):
            job["hs"].writes.append(self.writers.submit(flowio.write_tif, p["flow_file"], flow[k]))
            self.written.append(p["flow_file"])
            self.stats["bytes_written"] += flow[k].nbytes
            if p["warp_file"] and not os.path.isfile(p["warp_file"]):
                job["hs"].writes.append(self.writers.submit(flowio.write_tif, p["warp_file"], warped[k]))
                self.written.append(p["warp_file"])
                self.stats["bytes_written"] += warped[k].nbytes
        self.stats["batches"] += 1
        self.stats["pairs"] += len(job["batch"])

    def submit(self, frame_paths, batch, want_warp):
        nxt = self._start_read(frame_paths, batch, want_warp)      # batch i+1 starts reading ...
        if self.reading is not None:
            self._submit(*self.reading)                            # ... while batch i goes to the GPU
        self.reading = nxt

    def drain(self):
        if self.reading is not None:
            self._submit(*self.reading)
            self.reading = None
        self._retire(self.nsub & 1)                                # older slot first
        self._retire((self.nsub & 1) ^ 1)
        for hs in self.sets:
            for f in hs.writes:
                f.result()
            hs.writes = []
        self.stats["pinned_bytes"] = sum(hs.pinned_bytes() for hs in self.sets)
        return self.written

    def close(self):
        self.readers.shutdown()
        self.writers.shutdown()


def precompute_dataset(videos, flow_root, warp_root=None, patch_depth=2, future_patch_depth=0, rank=0, world=1,
                       compute=None, gather=None, max_pairs_per_batch=64, readers=None, writers=None, stats=None,
                       bridge=None):
    """``videos``: list of (name, [frame paths]).  Rank ``rank`` of ``world`` handles its round-robin shard.

    With ``compute=None`` the batches go through the GPU pipeline (upload / solve / download + file writing overlap
    across batches); an injected ``compute(frames, src, tgt, want_warp)`` is called synchronously per batch.
    ``gather`` (optional) receives this rank's list of written files and returns the lists of all ranks -- with
    torch.distributed that is ``all_gather_object``; it is the only communication of the whole job."""
    mine = shard(videos, rank, world)
    written = []
    pipe = _PipelinedGPU(readers, writers, bridge) if compute is None else None
    for name, paths in mine:
        flow_dir = os.path.join(flow_root, name)
        warp_dir = os.path.join(warp_root, name) if warp_root else None
        if pipe is None:
            written += precompute_video(paths, flow_dir, warp_dir, patch_depth, future_patch_depth, compute, max_pairs_per_batch)
            continue
        todo = plan_video(paths, flow_dir, warp_dir, patch_depth, future_patch_depth)
        if todo:
            os.makedirs(flow_dir, exist_ok=True)
            if warp_dir:
                os.makedirs(warp_dir, exist_ok=True)
            todo = _split_cached(todo, paths, written)
        for b0 in range(0, len(todo), max_pairs_per_batch):
            pipe.submit(paths, todo[b0:b0 + max_pairs_per_batch], warp_dir is not None)
    if pipe is not None:
        written += pipe.drain()
        pipe.close()
        if stats is not None:
            stats.update(pipe.stats)
    if gather is not None:
        return [f for part in gather(written) for f in part]
    return written


def torch_gather(written):
    """all_gather_object over the default process group (host-side gather of file names; no tensor collective)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, written)
    return out


def list_videos(noisy_root):
    """(name, sorted frame paths) for every sub-directory of ``noisy_root`` (library.py:93-115 file ordering)."""
    exts = ("tiff", "tif", "png", "jpg", "jpeg", "npy")
    videos = []
    for name in sorted(os.listdir(noisy_root)):
        d = os.path.join(noisy_root, name)
        if not os.path.isdir(d):
            continue
        files = sorted(os.listdir(d))
        for ext in exts:
            sel = [f for f in files if f.lower().endswith("." + ext)]
            if sel:
                videos.append((name, [os.path.join(d, f) for f in sel]))
                break
    return videos


def main(argv=None):
    """One process per GPU:  torchrun --nproc-per-node N -m rvdd_release_b200.precompute --noisy <dir> --flow <dir>"""
    import argparse
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--noisy", required=True, help="directory with one sub-directory of frames per video")
    ap.add_argument("--flow", required=True, help="flow cache root (<dataroot>/<flowFolder>/.../noisyinputs)")
    ap.add_argument("--warped", default=None)
    ap.add_argument("--patch-depth", type=int, default=2)
    ap.add_argument("--future-patch-depth", type=int, default=0)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    from . import hostbind
    hostbind.bind_to_gpu(int(os.environ.get("LOCAL_RANK", 0)))     # pinned staging buffers on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("gloo")
    files = precompute_dataset(list_videos(args.noisy), args.flow, args.warped, args.patch_depth, args.future_patch_depth,
                               rank, world, gather=torch_gather if world > 1 else None)
    if rank == 0:
        print("wrote %d files" % len(files))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
