"""Config 4 in its own shape: offline flow precompute of a REDS-shaped synthetic training set THROUGH FILES, sharded by
sequence over the ranks (one process per GPU).

    python profiles/bench_precompute.py --seqs-per-gpu 12                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/bench_precompute.py --seqs-per-gpu 12

Every sequence is 90 frames of 640x360x4 packed raw (REDS 1280x720 Bayer frames packed, ISO 3200 noise), stored as float32
TIFFs the way `iio.write` stores them; `precompute_dataset` (data/base_dataset.py:134-249 equivalent, patch_depth 2,
future_patch_depth 1 -> 178 pairs per sequence) reads them, computes the flows and writes one (h, w, 2) TIFF per pair.
The dataset (240 sequences in the reference's REDS split) is scaled to `--seqs-per-gpu` per GPU so the run takes seconds;
the work per sequence is the real one.  Timed: wall clock of the slowest rank around `precompute_dataset`, files included.
Beside it, the same batches through the same staging slots WITHOUT files (frames already in pinned memory, flows left in
pinned memory) -- the no-I/O ceiling of the pipeline -- and bare file read / write rates of the storage used.
Prints one JSON line (rank 0)."""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from rvdd_release_b200 import bridge, flowio, hostbind, precompute, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs-per-gpu", type=int, default=12)
    ap.add_argument("--frames", type=int, default=90)
    ap.add_argument("--h", type=int, default=360)
    ap.add_argument("--w", type=int, default=640)
    ap.add_argument("--future", type=int, default=1)
    ap.add_argument("--root", default="/dev/shm/rvdd_config4")
    ap.add_argument("--readers", type=int, default=0)
    ap.add_argument("--writers", type=int, default=0)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    binding = hostbind.bind_to_gpu(local)
    if world > 1:
        dist.init_process_group("gloo")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- the synthetic dataset (not timed): rank r writes the sequences it will not necessarily process itself
    noisy = os.path.join(args.root, "noisy")
    flow_root = os.path.join(args.root, "flow", "tvl1", "noisyinputs")
    nseq = args.seqs_per_gpu * world
    if rank == 0 and os.path.isdir(args.root):
        shutil.rmtree(args.root)
    barrier()
    t0 = time.perf_counter()
    for s in range(rank, nseq, world):
        d = os.path.join(noisy, "%03d" % s)
        os.makedirs(d, exist_ok=True)
        seq = synth.sequence(args.frames, args.h, args.w, "iso3200", device="cuda", noise_seed=977 * s).cpu().numpy()
        for f in range(args.frames):
            flowio.write_tif(os.path.join(d, "%08d.tif" % f), seq[f])
    barrier()
    gen_s = time.perf_counter() - t0
    videos = precompute.list_videos(noisy)
    assert len(videos) == nseq
    npairs_seq = len(precompute.video_pairs(args.frames, 2, args.future))

    # ---- bare storage rates with this rank's share of the files (page cache is warm for reads: the files were just
    # written; that is also the state the precompute sees)
    mine = precompute.shard(videos, rank, world)
    buf = np.empty((args.h, args.w, 4), np.float32)
    t0 = time.perf_counter()
    nread = 0
    for _, paths in mine[:2]:
        for p in paths:
            flowio.read_tif_into(p, buf)
            nread += buf.nbytes
    read_gbs = nread / (time.perf_counter() - t0) / 1e9
    tmpd = os.path.join(args.root, "scratch%d" % rank)
    os.makedirs(tmpd, exist_ok=True)
    fl = np.random.rand(args.h, args.w, 2).astype(np.float32)
    t0 = time.perf_counter()
    for i in range(100):
        flowio.write_tif(os.path.join(tmpd, "%d.tif" % i), fl)
    write_gbs = 100 * fl.nbytes / (time.perf_counter() - t0) / 1e9
    shutil.rmtree(tmpd)

    br = bridge.default_bridge()
    # warm-up (kernels, workspace, pinned staging) on one video into a scratch directory
    warm = os.path.join(args.root, "warm%d" % rank)
    precompute.precompute_dataset(mine[:1], warm, None, 2, args.future, max_pairs_per_batch=args.batch,
                                  readers=args.readers or None, writers=args.writers or None)
    shutil.rmtree(warm)

    # ---- timed: the real thing, files in, files out
    stats = {}
    barrier()
    t0 = time.perf_counter()
    files = precompute.precompute_dataset(videos, flow_root, None, 2, args.future, rank, world, max_pairs_per_batch=args.batch,
                                          readers=args.readers or None, writers=args.writers or None, stats=stats)
    torch.cuda.synchronize()
    mine_s = time.perf_counter() - t0
    barrier()
    total_s = max_over_ranks(mine_s)
    nfiles = int(sum_over_ranks(len(files)))
    assert nfiles == nseq * npairs_seq, (nfiles, nseq * npairs_seq)

    # ---- the same batches without files: pinned frames -> staging slots -> pinned flows
    plan = precompute.plan_video(mine[0][1], os.path.join(args.root, "nowhere"), None, 2, args.future)
    batches = [plan[i:i + args.batch] for i in range(0, len(plan), args.batch)]
    jobs = []
    for b in batches:
        used = sorted({p["src"] for p in b} | {p["tgt"] for p in b})
        loc = {f: i for i, f in enumerate(used)}
        fr = torch.empty((len(used), args.h, args.w, 4), dtype=torch.float32).pin_memory()
        for i, f in enumerate(used):
            flowio.read_tif_into(mine[0][1][f], fr[i].numpy())
        jobs.append((fr, [loc[p["src"]] for p in b], [loc[p["tgt"]] for p in b],
                     torch.empty((len(b), args.h, args.w, 2), dtype=torch.float32).pin_memory()))
    reps = len(mine)
    barrier()
    t0 = time.perf_counter()
    k = 0
    for _ in range(reps):
        for fr, s_, t_, fo in jobs:
            br.wait_host(k & 1)
            br.submit_host(k & 1, fr, s_, t_, fo)
            k += 1
    br.wait_host(0)
    br.wait_host(1)
    noio_s = max_over_ranks(time.perf_counter() - t0)
    noio_pairs = world * reps * len(plan)

    # ---- the storage + host-memory ceiling: the same file reads (into pinned memory) and writes (from pinned memory) with the
    # same thread pools on all ranks at once, no GPU work, no host<->device copies
    from concurrent.futures import ThreadPoolExecutor
    ncpu = max(1, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
    nthr = args.readers or max(2, min(8, ncpu // 2))
    pool_r, pool_w = ThreadPoolExecutor(nthr), ThreadPoolExecutor(args.writers or nthr)
    io_dir = os.path.join(args.root, "io_only%d" % rank)
    os.makedirs(io_dir, exist_ok=True)
    pin_fr = torch.empty((args.frames, args.h, args.w, 4), dtype=torch.float32).pin_memory().numpy()
    pin_fl = torch.rand((npairs_seq, args.h, args.w, 2), dtype=torch.float32).pin_memory().numpy()
    barrier()
    t0 = time.perf_counter()
    futs = []
    for vi, (_, paths) in enumerate(mine):
        futs += [pool_r.submit(flowio.read_tif_into, p, pin_fr[i]) for i, p in enumerate(paths)]
        futs += [pool_w.submit(flowio.write_tif, os.path.join(io_dir, "%d_%d.tif" % (vi, k)), pin_fl[k]) for k in range(npairs_seq)]
        if vi >= 1:                                      # keep about two videos of requests in flight, like the pipeline
            for f in futs[:args.frames + npairs_seq]:
                f.result()
            futs = futs[args.frames + npairs_seq:]
    for f in futs:
        f.result()
    io_s = max_over_ranks(time.perf_counter() - t0)
    io_pairs = world * len(mine) * npairs_seq
    pool_r.shutdown()
    pool_w.shutdown()
    shutil.rmtree(io_dir)

    # ---- spot check: a file of this rank against a direct device computation
    name, paths = mine[-1]
    fr = torch.from_numpy(np.stack([flowio.read_image(paths[i]) for i in (4, 5)])).cuda()
    direct = br.tvl1_flow(br.gray(fr), [0, 1], [1, 0], check=True).cpu().numpy()
    f_past = flowio.read_tif(os.path.join(flow_root, name, "%08d_%08d.tif" % (4, 5)))
    f_fut = flowio.read_tif(os.path.join(flow_root, name, "%08d_%08d.tif" % (5, 4)))
    ok = bool(np.array_equal(f_past.transpose(2, 0, 1), direct[0]) and np.array_equal(f_fut.transpose(2, 0, 1), direct[1]))
    ok = bool(sum_over_ranks(0.0 if ok else 1.0) == 0.0)

    if rank == 0:
        pairs = nseq * npairs_seq
        print(json.dumps({
            "what": "config 4: offline flow precompute through files (precompute_dataset)", "n_gpus": world,
            "sequences": nseq, "frames_per_sequence": args.frames, "frame": [args.h, args.w, 4], "pairs": pairs,
            "pairs_per_s_files_included": pairs / total_s, "seconds": total_s,
            "pairs_per_s_no_io_same_batches": noio_pairs / noio_s,
            "files_vs_no_io": (pairs / total_s) / (noio_pairs / noio_s),
            "pairs_per_s_io_only_ceiling": io_pairs / io_s,
            "files_vs_min_of_ceilings": (pairs / total_s) / min(noio_pairs / noio_s, io_pairs / io_s),
            "storage": args.root, "bare_read_GBps_1thread": read_gbs, "bare_write_GBps_1thread": write_gbs,
            "read_GBps_achieved_per_rank": stats.get("bytes_read", 0) / mine_s / 1e9,
            "write_GBps_achieved_per_rank": stats.get("bytes_written", 0) / mine_s / 1e9,
            "rank0_waits_s": {k: round(stats.get(k, 0.0), 3) for k in ("wait_read_s", "wait_gpu_s", "wait_write_s")},
            "rank0_seconds": mine_s,
            "readers": args.readers or "auto", "writers": args.writers or "auto", "batch_pairs": args.batch,
            "pinned_MB_per_rank": stats.get("pinned_bytes", 0) / 1e6, "host_cpus": os.cpu_count(),
            "host_binding": binding, "dataset_generation_s": gen_s, "files_bit_equal_to_direct_compute": ok,
            "scaled_from": "240 sequences x 90 frames (BASELINE.json configs[3]); per-sequence work unchanged"}))
    if not args.keep:
        barrier()
        if rank == 0:
            shutil.rmtree(args.root, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
