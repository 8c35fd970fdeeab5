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
    """The pairs of one video that still need work -> list of dicts(src, tgt, flow_file, warp_file)."""
    todo = []
    for s, t in video_pairs(len(frame_paths), patch_depth, future_patch_depth):
        ff = warpedimagefile(flow_dir, _code(frame_paths[s]), _code(frame_paths[t]))
        wf = warpedimagefile(warp_dir, _code(frame_paths[s]), _code(frame_paths[t])) if warp_dir else None
        if not os.path.isfile(ff) or (wf and not os.path.isfile(wf)):
            todo.append(dict(src=s, tgt=t, flow_file=ff, warp_file=wf))
    return todo


def gpu_compute(frames, src, tgt, want_warp):
    """Default compute function: one batched call into libBridge.so with host buffers."""
    from . import bridge
    flow, warped, _ = bridge.default_bridge().flow_and_warp_host(frames, src, tgt, want_warp=want_warp)
    return flow.numpy(), (warped.numpy() if want_warp else None)


def precompute_video(frame_paths, flow_dir, warp_dir=None, patch_depth=2, future_patch_depth=0, compute=gpu_compute,
                     max_pairs_per_batch=64):
    """Create the missing flow (and optionally warped) files of one video.  Returns the files written."""
    todo = plan_video(frame_paths, flow_dir, warp_dir, patch_depth, future_patch_depth)
    if not todo:
        return []
    os.makedirs(flow_dir, exist_ok=True)
    if warp_dir:
        os.makedirs(warp_dir, exist_ok=True)
    written = []
    for b0 in range(0, len(todo), max_pairs_per_batch):
        batch = todo[b0:b0 + max_pairs_per_batch]
        used = sorted({p["src"] for p in batch} | {p["tgt"] for p in batch})
        local = {f: i for i, f in enumerate(used)}
        frames = np.stack([flowio.read_image(frame_paths[f]).astype(np.float32) for f in used])   # base_dataset.py:159,174
        src = [local[p["src"]] for p in batch]
        tgt = [local[p["tgt"]] for p in batch]
        flow, warped = compute(frames, src, tgt, warp_dir is not None)
        for k, p in enumerate(batch):
            if not os.path.isfile(p["flow_file"]):
                flowio.write_tif(p["flow_file"], flow[k])                                     # base_dataset.py:180
                written.append(p["flow_file"])
            if p["warp_file"] and not os.path.isfile(p["warp_file"]):
                flowio.write_tif(p["warp_file"], warped[k])                                   # base_dataset.py:189
                written.append(p["warp_file"])
    return written


class _PipelinedGPU:
    """Feeds batches to libBridge.so through its two staging slots (rvdd_flow_and_warp_host_submit / _wait): while
    one batch is on the GPU the previous one is written to disk and the next one is read and uploaded."""

    def __init__(self):
        import torch
        from . import bridge
        self.torch, self.br = torch, bridge.default_bridge()
        self.jobs = [None, None]
        self.k = 0
        self.written = []

    def _buffers(self, slot, nfr, h, w, c, npairs, want_warp):
        t = self.torch
        key = (nfr, h, w, c, npairs, want_warp)
        cache = getattr(self, "_cache", None) or {}
        self._cache = cache
        if (slot, key) not in cache:
            cache[(slot, key)] = (t.empty((nfr, h, w, c), dtype=t.float32).pin_memory(),
                                  t.empty((npairs, h, w, 2), dtype=t.float32).pin_memory(),
                                  t.empty((npairs, h, w, c), dtype=t.float32).pin_memory() if want_warp else None)
        return cache[(slot, key)]

    def _finish(self, slot):
        job = self.jobs[slot]
        if job is None:
            return
        self.br.wait_host(slot)
        batch, flow, warped = job
        for k, p in enumerate(batch):
            if not os.path.isfile(p["flow_file"]):
                flowio.write_tif(p["flow_file"], flow[k].numpy())
                self.written.append(p["flow_file"])
            if p["warp_file"] and not os.path.isfile(p["warp_file"]):
                flowio.write_tif(p["warp_file"], warped[k].numpy())
                self.written.append(p["warp_file"])
        self.jobs[slot] = None

    def submit(self, frame_paths, batch, want_warp):
        slot = self.k & 1
        self.k += 1
        self._finish(slot)                                  # this slot's previous batch: wait + write its files
        used = sorted({p["src"] for p in batch} | {p["tgt"] for p in batch})
        local = {f: i for i, f in enumerate(used)}
        first = flowio.read_image(frame_paths[used[0]]).astype(np.float32)
        h, w, c = first.shape
        frames, flow, warped = self._buffers(slot, len(used), h, w, c, len(batch), want_warp)
        frames[0].copy_(self.torch.from_numpy(first))
        for i, f in enumerate(used[1:], 1):
            frames[i].copy_(self.torch.from_numpy(flowio.read_image(frame_paths[f]).astype(np.float32)))
        self.br.submit_host(slot, frames, [local[p["src"]] for p in batch], [local[p["tgt"]] for p in batch], flow, warped)
        self.jobs[slot] = (batch, flow, warped)

    def drain(self):
        self._finish(0)
        self._finish(1)
        return self.written


def precompute_dataset(videos, flow_root, warp_root=None, patch_depth=2, future_patch_depth=0, rank=0, world=1,
                       compute=None, gather=None, max_pairs_per_batch=64):
    """``videos``: list of (name, [frame paths]).  Rank ``rank`` of ``world`` handles its round-robin shard.

    With ``compute=None`` the batches go through the GPU pipeline (upload / solve / download + file writing overlap
    across batches); an injected ``compute(frames, src, tgt, want_warp)`` is called synchronously per batch.
    ``gather`` (optional) receives this rank's list of written files and returns the lists of all ranks -- with
    torch.distributed that is ``all_gather_object``; it is the only communication of the whole job."""
    mine = shard(videos, rank, world)
    written = []
    pipe = _PipelinedGPU() if compute is None else None
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
        for b0 in range(0, len(todo), max_pairs_per_batch):
            pipe.submit(paths, todo[b0:b0 + max_pairs_per_batch], warp_dir is not None)
    if pipe is not None:
        written += pipe.drain()
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
