"""Batched multi-sequence recurrent inference: the alignment path of the reference's validation loop
(validate.py:75-106 -> models/recurrent_model.py:102-129, :233-345) with the SEQUENCES as the batch dimension.

The reference denoises one sequence at a time, frame by frame (batch 1, validate.py:46-48), after its dataset constructor
has computed every flow on the CPU one pair at a time (data/infer4rec_dataset.py:127-128 -> base_dataset.py:134-249).  A
video is serial in time (the previous denoised frame and its features feed the next frame) but videos are independent, so
here S sequences advance together:

    all flows of all sequences      ONE batched solver call (past t-1 -> t and future t+1 -> t pairs)
    all frames demosaicked          one launch (HamiltonAdam, recurrent_model.py:126)
    per time step t                 FrameAligner.step with B = S: warps of the previous denoised frames, their feature maps
                                    and the future frames straight into the network input (half-resolution flows, the x2
                                    upsampling of recurrent_model.py:129 fused), then the denoiser on the batch

The denoiser is NOT part of this package (cuDNN networks of the reference, networks/*.py): it is any callable
`den = denoiser(netinput)` or `den, feat = denoiser(netinput, featinput)`, e.g. a TorchScript export of a shipped
checkpoint.  Multi-GPU: one process per GPU, sequences sharded round-robin over the ranks, no collective on the data path;
`main` gathers the per-sequence PSNR on the host at the end.
"""
import os

import numpy as np
import torch

from . import bridge as _bridge
from .hamilton_adam import HamiltonAdam
from .recurrent_align import FrameAligner


def sequence_pairs(S, T, future_depth):
    """(src, tgt) frame indices into the flattened [S * T] frame list: past pairs t-1 -> t for t = 1 .. T-1-fD, then (with
    future_depth 1) future pairs t+1 -> t for the same t -- per sequence the order of data/infer4rec_dataset.py:198-202."""
    last = T - future_depth
    src, tgt = [], []
    for s in range(S):
        for t in range(1, last):
            src.append(s * T + t - 1)
            tgt.append(s * T + t)
    for s in range(S):
        for t in range(1, last):
            for b in range(future_depth):
                src.append(s * T + t + 1 + b)
                tgt.append(s * T + t)
    return np.asarray(src, np.int32), np.asarray(tgt, np.int32)


def compute_all_flows(frames, future_depth=0, max_pairs_per_call=4096):
    """frames: CUDA float32 [S, T, h, w, c] packed frames (raw units; the flow is invariant to their affine scaling only
    through the joint normalisation, so pass what the dataset holds) -> (past [S, T-1-fD, 2, h, w], future [S, T-1-fD, fD, 2,
    h, w] or None), flows at the packed resolution, source -> target t."""
    br = _bridge.default_bridge()
    S, T, h, w, c = frames.shape
    gray = br.gray(frames.reshape(S * T, h, w, c))
    src, tgt = sequence_pairs(S, T, future_depth)
    out = torch.empty((len(src), 2, h, w), dtype=torch.float32, device=frames.device)
    for k0 in range(0, len(src), max_pairs_per_call):
        k1 = min(len(src), k0 + max_pairs_per_call)
        out[k0:k1] = br.tvl1_flow(gray, src[k0:k1], tgt[k0:k1])
    br.check(frames.device)
    n = T - 1 - future_depth
    past = out[:S * n].view(S, n, 2, h, w)
    fut = out[S * n:].view(S, n, future_depth, 2, h, w) if future_depth else None
    return past, fut


def _call_denoiser(denoiser, netinput, featinput, sub_batch):
    B = netinput.shape[0]
    sub = sub_batch or B
    dens, feats = [], []
    for b0 in range(0, B, sub):
        if featinput is not None:
            d, f = denoiser(netinput[b0:b0 + sub], featinput[b0:b0 + sub])
            feats.append(f)
        else:
            d = denoiser(netinput[b0:b0 + sub])
        dens.append(d)
    den = dens[0] if len(dens) == 1 else torch.cat(dens, 0)
    feat = None if featinput is None else (feats[0] if len(feats) == 1 else torch.cat(feats, 0))
    return den, feat


def run_sequences(frames, denoiser, future_depth=0, feature_channels=0, bit_depth=12, pattern="gbrg", denoiser_batch=None,
                  on_frame=None, flows=None):
    """Recurrent inference of S sequences in lock step.

    frames   : CUDA float32 [S, T, h, w, 4] packed raw frames in [0, 2^bit_depth - 1] (data/infer4rec_dataset.py:195-218
               divides by 2^bit_depth - 1 and maps to [-1, 1]);
    denoiser : callable, see the module docstring;  denoiser_batch: run it on this many sequences at a time (memory);
    on_frame : optional callback(t, denoised [S, 3, 2h, 2w]) per time step (PSNR, writing frames, ...); without it the
               denoised frames are returned as a list;
    flows    : optional precomputed (past, future) as returned by compute_all_flows (e.g. read from the flow cache).
    Returns (list of denoised batches or None, timings dict with CUDA-event milliseconds: flow, demosaic, align, denoise)."""
    S, T, h, w, c = frames.shape
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    e = [ev() for _ in range(3)]
    e[0].record()
    past, fut = flows if flows is not None else compute_all_flows(frames, future_depth)
    e[1].record()
    ha = HamiltonAdam(pattern)
    maxv = float(2 ** bit_depth - 1)
    scale = torch.tensor(maxv, dtype=frames.dtype, device=frames.device)       # tensor divisor: a true division on CUDA too
    n = ha((2.0 * (frames.reshape(S * T, h, w, c) / scale) - 1.0).permute(0, 3, 1, 2).contiguous())
    n = n.view(S, T, 3, 2 * h, 2 * w)
    e[2].record()
    al = FrameAligner(depth=1, future_depth=future_depth, feature_channels=feature_channels, pattern=pattern)
    al.reset(n[:, 0])
    outs = [] if on_frame is None else None
    t_align = t_den = 0.0
    marks = []
    for t in range(1, T - future_depth):
        a, b, d = ev(), ev(), ev()
        a.record()
        netinput, featinput = al.step(n[:, t], past[:, t - 1], [n[:, t + 1 + k] for k in range(future_depth)],
                                      [fut[:, t - 1, k] for k in range(future_depth)])
        b.record()
        den, feat = _call_denoiser(denoiser, netinput, featinput, denoiser_batch)
        al.update(den, feat)
        d.record()
        marks.append((a, b, d))
        if on_frame is not None:
            on_frame(t, den)
        else:
            outs.append(den)
    torch.cuda.synchronize(frames.device)
    for a, b, d in marks:
        t_align += a.elapsed_time(b)
        t_den += b.elapsed_time(d)
    timings = dict(flow_ms=e[0].elapsed_time(e[1]), demosaic_ms=e[1].elapsed_time(e[2]), align_ms=t_align, denoise_ms=t_den,
                   frames=S * (T - 1 - future_depth), pairs=int(past.shape[0] * past.shape[1] * (1 + future_depth)))
    return outs, timings


def psnr(a, b, max_val=2.0):
    """util/util.py:9-20 per batch element -> tensor [B]"""
    mse = ((a - b) ** 2).flatten(1).mean(1)
    return 10.0 * torch.log10(max_val * max_val / mse)


def main(argv=None):
    """One process per GPU:  torchrun --nproc-per-node N -m rvdd_release_b200.infer --denoiser net.pt --sequences 64 ...
    Synthetic sequences (synth.exact_sequence, one noise realisation per sequence) unless --frames-npy points at an
    [S, T, h, w, 4] array; prints one JSON line with frames/s and the time split."""
    import argparse
    import json
    import time
    import torch.distributed as dist
    from . import hostbind, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--denoiser", required=True, help="TorchScript denoiser (netinput[, featinput]) -> denoised[, features]")
    ap.add_argument("--sequences", type=int, default=8, help="total number of sequences (sharded over the ranks)")
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--iso", default="iso3200")
    ap.add_argument("--future-depth", type=int, default=1)
    ap.add_argument("--feature-channels", type=int, default=48)
    ap.add_argument("--denoiser-batch", type=int, default=1)
    ap.add_argument("--frames-npy", default=None)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    hostbind.bind_to_gpu(local)
    if world > 1:
        dist.init_process_group("gloo")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    mine = [s for s in range(args.sequences) if s % world == rank]
    if args.frames_npy:
        frames = torch.from_numpy(np.load(args.frames_npy)[mine]).cuda()
    else:
        frames = torch.stack([synth.exact_sequence(args.frames, args.h, args.w, args.iso, device="cuda", noise_seed=s)
                              for s in mine])
    net = torch.jit.load(args.denoiser, map_location="cuda").eval()
    cfg = synth.ISO[args.iso]
    ha = HamiltonAdam("gbrg")
    gts = None
    if not args.frames_npy:
        gts = [(2.0 * ha.pack_in_one((cfg["lo"] + synth.exact_clean_frame(t, args.h, args.w, device="cuda") * (cfg["hi"] - cfg["lo"]))
                                     .float().permute(2, 0, 1)[None]) / 4095.0 - 1.0)[:, None].repeat(1, 3, 1, 1)
               for t in range(args.frames)]
    acc = []

    def on_frame(t, den):
        if gts is not None:
            acc.append(psnr(den, gts[t].expand_as(den)).cpu())

    with torch.no_grad():
        run_sequences(frames, net, args.future_depth, args.feature_channels, denoiser_batch=args.denoiser_batch)   # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            acc.clear()
            _, tm = run_sequences(frames, net, args.future_depth, args.feature_channels, denoiser_batch=args.denoiser_batch,
                                  on_frame=on_frame)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
    per_seq = torch.stack(acc, 1).mean(1).tolist() if acc else []
    info = dict(rank=rank, seconds=dt, sequences=len(mine), psnr_per_sequence=per_seq, **tm)
    allinfo = [info]
    if world > 1:
        allinfo = [None] * world
        dist.all_gather_object(allinfo, info)            # host-side gather of scalars: the only communication
    if rank == 0:
        worst = max(i["seconds"] for i in allinfo)
        nfr = sum(i["frames"] for i in allinfo)
        print(json.dumps({
            "what": "batched multi-sequence recurrent inference (config 5 shape)", "n_gpus": world,
            "sequences": args.sequences, "frames_per_sequence": args.frames, "packed_frame": [args.h, args.w, 4],
            "network_resolution": [2 * args.h, 2 * args.w], "denoised_frames": nfr, "seconds": worst,
            "denoised_frames_per_s": nfr / worst, "flow_pairs": sum(i["pairs"] for i in allinfo),
            "per_rank_ms": {k: max(i[k] for i in allinfo) for k in ("flow_ms", "demosaic_ms", "align_ms", "denoise_ms")},
            "alignment_share_of_time": max((i["flow_ms"] + i["demosaic_ms"] + i["align_ms"]) / (1e3 * i["seconds"]) for i in allinfo),
            "mean_psnr": float(np.mean([p for i in allinfo for p in i["psnr_per_sequence"]])) if per_seq else None,
            "denoiser": os.path.basename(args.denoiser), "denoiser_batch": args.denoiser_batch}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
