#!/usr/bin/env python
"""Headline benchmark: ResNet-50 (reference symbol, 25.5 M params) synthetic-data training throughput,
images/sec over the whole job, bf16, device-timed, max over ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode allreduce|ps_sync|ps_dedicated|nccl|nccl_ps]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...   -> the unmodified reference (unavailable offline)

Reference analogue: `train_imagenet.py --benchmark 1 --kv-store device|dist_sync_device`
(MX/example/image-classification/README.md:253-258) and the Horovod `train_synthetic.sh` (README.md:149-163).
Published comparator: ResNet-50 training 363.69 img/s on one V100 (MX/docs/faq/perf.md:219).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

BASELINE_IMG_S_PER_GPU = 363.69  # BASELINE.md row 13 (V100, batch 128, fp32) - the reference's best published ResNet-50 training number


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "nccl", "torch-stock"],
                    help="ours | reference (the unmodified reference: unavailable offline) | nccl (ours with NCCL collectives) | "
                         "torch-stock (stock nn.Module ResNet-50 + torch.optim.SGD + DDP, baseline/torch_stock.py: no repo code)")
    ap.add_argument("--stock", type=int, default=int(os.environ.get("DLCFN_BENCH_STOCK", "1")),
                    help="after the own measurement, run the stock-PyTorch arm (baseline/torch_stock.py) in a child process and add "
                         "`vs_stock` to the JSON line (ResNet only).  0 = never, 1 = single-GPU runs only (default), 2 = also under "
                         "torchrun (one child per rank with its own process group); multi-GPU stock numbers are otherwise taken with "
                         "`--impl torch-stock` (profiles/r2/)")
    ap.add_argument("--mode", default=None, help="allreduce (default) | ps_sync | ps_dedicated | nccl | nccl_ps | hier (multi-node: fused node-local + NCCL rails)")
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "bert-large", "bert-base"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 256 images / 16 sequences)")
    ap.add_argument("--seq-len", type=int, default=512)
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--num-layers", type=int, default=50)
    ap.add_argument("--bucket-mb", type=float, default=16.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dry-cpu", action="store_true",
                    help="exercise the whole script on the host with tiny shapes (fp32, wall-clock 'events', no H2D): a check "
                         "of this file's logic on a box without a GPU - the line it prints is marked and is NOT a measurement")
    ap.add_argument("--device-batches", type=int, default=4,
                    help="distinct device-resident input batches rotated by the device-timed loop (the same synthetic images the "
                         "end-to-end loop streams from the host); 1 = train on one fixed batch")
    ap.add_argument("--cuda-graph", action=argparse.BooleanOptionalAction, default=os.environ.get("DLCFN_BENCH_GRAPH", "1") != "0",
                    help="replay the training step (forward, backward, fused collectives, optimizer: every kernel of the eager step) as "
                         "one CUDA graph captured through the public API `trainer.capture` (parallel/graph.py); --no-cuda-graph = eager")
    ap.add_argument("--conv1x1", default=None, choices=["tcgen05", "cudnn", "auto"],
                    help="backend of the 1x1 convolutions (default: ops.conv default / DLCFN_CONV1X1)")
    args, unknown = ap.parse_known_args()
    if unknown:      # a harness flag this script does not know must not cost the measurement
        print(f"[bench] ignoring unknown arguments: {unknown}", file=sys.stderr, flush=True)
    return args


def reference_arm(args):
    # The reference is CloudFormation JSON + a Python-2/boto agent + git submodules of MXNet 1.3 / TF 1.x;
    # `pip install /root/reference` finds no setup.py/pyproject, and the vendored MXNet python package needs a
    # libmxnet.so built with CUDA<=10 / cuDNN 7 (no sm_100 support).  Outcome recorded in DESIGN.md.
    print(json.dumps({"impl": "reference",
                      "unavailable": "reference is not pip-installable offline (no setup.py; vendored MXNet 1.3 needs "
                                     "libmxnet.so built with CUDA<=10/cuDNN7, no sm_100) - see DESIGN.md"}))
    return 0


class _HostEvent:
    """--dry-cpu stand-in for torch.cuda.Event (wall clock)."""

    def __init__(self, enable_timing: bool = False):
        self.t = 0.0

    def record(self, *_a):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other) -> float:
        return (other.t - self.t) * 1e3


class _HostPrefetcher:
    """--dry-cpu stand-in for utils.data.DevicePrefetcher: same interface, normalisation through torch on the host."""

    def __init__(self, source, device):
        self.source = source

    def next(self):
        import torch
        img, lab = self.source.next()
        return ((img.float() / 255.0 - 0.45) / 0.225).permute(0, 3, 1, 2), lab


def stock_argv(args):
    here = os.path.dirname(os.path.abspath(__file__))
    argv = [sys.executable, os.path.join(here, "baseline", "torch_stock.py"), "--gpus", str(args.gpus), "--steps", str(args.steps),
            "--warmup", str(args.warmup), "--batch", str(args.batch or 256), "--image-size", str(args.image_size),
            "--device-batches", str(args.device_batches)]
    if args.no_e2e:
        argv.append("--no-e2e")
    return argv


def run_stock_child(args, rank):
    """The stock-PyTorch arm in a child process of every rank (its own NCCL group on MASTER_PORT + 17): nothing of this
    repository is imported there.  Returns rank 0's parsed JSON record, or None."""
    import subprocess
    env = dict(os.environ)
    if "MASTER_PORT" in env:
        env["MASTER_PORT"] = str(int(env["MASTER_PORT"]) + 17)
    for k in [k for k in env if k.startswith("TORCHELASTIC_")]:
        env.pop(k)        # the child group makes its own TCPStore on MASTER_PORT + 17 (the agent's store lives on MASTER_PORT)
    try:
        r = subprocess.run(stock_argv(args), env=env, capture_output=True, text=True, timeout=300)
    except (subprocess.TimeoutExpired, OSError) as e:      # the headline must not depend on the comparison arm
        print(f"[bench] stock arm failed: {e}", file=sys.stderr, flush=True)
        return None
    if rank != 0:
        return None
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    print(f"[bench] stock arm printed no record (rc {r.returncode}): {r.stderr[-800:]}", file=sys.stderr, flush=True)
    return None


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    if args.impl == "torch-stock":
        os.execv(sys.executable, stock_argv(args))

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    dry = args.dry_cpu
    if not dry and not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device"}))
        return 1
    if dry:
        device = torch.device("cpu")
        Event, cuda_sync, pin = _HostEvent, (lambda _d=None: None), (lambda t: t)
        args.image_size, args.num_layers = min(args.image_size, 32), (20 if args.num_layers == 50 else args.num_layers)
        args.batch, args.seq_len = args.batch or 4, min(args.seq_len, 32)
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        Event, cuda_sync, pin = torch.cuda.Event, torch.cuda.synchronize, (lambda t: t.pin_memory())
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=device)

    from deeplearning_cfn_b200.models import resnet, to_b200_layout
    from deeplearning_cfn_b200.ops import native
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    from deeplearning_cfn_b200.utils.clocks import ClockSampler
    from deeplearning_cfn_b200.utils.data import DevicePrefetcher, SyntheticImageSource, synthetic_device_batch

    if not dry:
        native.lib()  # fail loudly if the sm_100a library is missing
    compute_dtype = torch.float32 if dry else torch.bfloat16
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)

    mode = args.mode or ("nccl" if args.impl == "nccl" else "allreduce")
    servers = workers = None
    if mode == "ps_dedicated":  # BASELINE config 3: 1 server + (N-1) workers
        mode, servers, workers = "ps_sync", [0], list(range(1, world))
    if world == 1:
        mode = "local"
    elif dry:
        mode = "nccl"          # torch.distributed all-reduce (gloo on the host); the fused kernels need peer HBM

    from deeplearning_cfn_b200.ops import conv as conv_ops
    if args.conv1x1:
        conv_ops.set_conv1x1_backend(args.conv1x1)
    is_bert = args.model.startswith("bert")
    if args.batch is None:
        args.batch = 16 if is_bert else 256
    if is_bert:
        from deeplearning_cfn_b200.models import bert_pretrain_loss, get_model
        args.no_e2e = True
        if dry:
            from deeplearning_cfn_b200.models import Bert
            model = Bert(hidden=64, layers=2, heads=4, ffn=128, max_seq=args.seq_len)
        else:
            model = get_model(args.model, max_seq=args.seq_len)
        model = to_b200_layout(model.to(device), compute_dtype).train()
        model.padded_logits = True     # the fused loss kernel masks the TMA padding columns of the vocabulary
        trainer = DataParallelTrainer(model, mode=mode, optimizer="adam", lr=1e-4, wd=0.01, beta1=0.9, beta2=0.999,
                                      bucket_bytes=int(args.bucket_mb * (1 << 20)), server_ranks=servers,
                                      worker_ranks=workers, loss_fn=bert_pretrain_loss)
        g = torch.Generator().manual_seed(0)
        npred = 76      # max_predictions_per_seq of the standard BERT pre-training recipe
        ids = torch.randint(0, 30522, (args.batch, args.seq_len), generator=g).to(device)
        x = (ids, torch.zeros_like(ids), torch.randint(0, args.seq_len, (args.batch, npred), generator=g).to(device))
        y = (torch.randint(0, 30522, (args.batch, npred), generator=g).to(device),
             torch.randint(0, 2, (args.batch,), generator=g).to(device))
        dev_batches = [(x, y)]
    else:
        model = to_b200_layout(resnet(args.num_layers, 1000, args.image_size).to(device), compute_dtype)
        model.train()
        trainer = DataParallelTrainer(model, mode=mode, optimizer="sgd_mom", lr=0.1, momentum=0.9, wd=1e-4,
                                      bucket_bytes=int(args.bucket_mb * (1 << 20)), server_ranks=servers,
                                      worker_ranks=workers)
        x, y = synthetic_device_batch(args.batch, args.image_size, 1000, device, dtype=compute_dtype)
        # The device-timed loop rotates the SAME synthetic uint8 image batches that the end-to-end loop streams from pinned host
        # memory (normalised once, resident on the device), so `value` and `e2e.value` differ by the input pipeline only.  A
        # single batch trained on for the whole run collapses the activations of a randomly initialised network within a few
        # steps, which round 1 measured as a 7 % shorter step than on fresh batches with identical kernels.
        if args.device_batches > 1:
            from deeplearning_cfn_b200.utils.data import IMAGENET_MEAN, IMAGENET_STD
            _src = SyntheticImageSource(args.batch, args.image_size, 1000, distinct=args.device_batches, pin=False)
            dev_batches = []
            for img, lab in zip(_src.images, _src.labels):
                if dry:
                    xb = ((img.float() / 255.0 - 0.45) / 0.225).permute(0, 3, 1, 2)
                else:
                    u8 = img.to(device)
                    nx = torch.empty(u8.shape, dtype=torch.bfloat16, device=device)
                    native.normalize_u8_nhwc(u8, nx, IMAGENET_MEAN, IMAGENET_STD)
                    xb = nx.permute(0, 3, 1, 2)            # NHWC storage viewed as NCHW = channels_last
                dev_batches.append((xb, lab.to(device)))
            del _src
        else:
            dev_batches = [(x, y)]
    n_workers = len(trainer.workers)
    is_worker = trainer.is_worker

    def sync_all():
        if world > 1:
            dist.barrier()
        cuda_sync(device)

    # ---------------------------------------------------------------- device-timed loop
    run_step = trainer.step
    graphed = None
    # ranks that only serve (dedicated parameter servers) run no forward / backward: their step cannot be captured next to
    # the workers' (measured at N = 8: cudaErrorStreamCaptureIsolation on the server rank), so that layout runs eagerly
    if args.cuda_graph and servers is not None:
        args.cuda_graph = False
        if rank == 0:
            print("[bench] dedicated servers: eager steps (no CUDA graph)", file=sys.stderr, flush=True)
    if args.cuda_graph:
        try:
            graphed = trainer.capture(x if is_worker else None, y if is_worker else None)
            run_step = graphed
        except Exception as e:     # the headline must not depend on the capture: a capture records without executing, so a rank
            graphed = None         # that falls back to eager steps stays in step with peers that replay their graphs
            print(f"[bench] rank {rank}: CUDA graph capture failed ({type(e).__name__}: {e}); eager steps", file=sys.stderr, flush=True)
            try:
                cuda_sync(device)
            except Exception:
                pass
        if rank == 0:
            print(f"[bench] {graphed!r}", file=sys.stderr, flush=True)
    nb = len(dev_batches)
    for i in range(args.warmup):
        run_step(*dev_batches[i % nb]) if is_worker else run_step()
    sync_all()
    launches0 = native.launch_count
    t_start, t_end = Event(enable_timing=True), Event(enable_timing=True)
    sampler = ClockSampler(gpu_index=local_rank) if rank == 0 else None
    if sampler:
        sampler.__enter__()
    t_start.record()
    step_loss = None
    marks = [Event(enable_timing=True) for _ in range(args.steps)]        # one event per step: min / median step time
    for i in range(args.steps):
        step_loss = run_step(*dev_batches[(args.warmup + i) % nb]) if is_worker else run_step()
        marks[i].record()
    t_end.record()
    sync_all()
    step_ms = sorted(a.elapsed_time(b) for a, b in zip([t_start] + marks[:-1], marks))
    if sampler:
        sampler.__exit__(None, None, None)
    ms = t_start.elapsed_time(t_end)
    # read after the timed region: the loss of the last timed step (a diverged / NaN run would not be a training step)
    timed_loss = float(step_loss.detach().float()) if (is_worker and torch.is_tensor(step_loss)) else None
    if timed_loss is not None and timed_loss != timed_loss or timed_loss in (float("inf"), float("-inf")):
        print(f"[bench] WARNING: loss of the last timed step is {timed_loss}", file=sys.stderr, flush=True)
        timed_loss = None             # keep the output line strict JSON (no NaN literal)
    launches = native.launch_count - launches0
    trainer.check()

    # ---------------------------------------------------------------- end-to-end loop (public API + H2D/D2H)
    e2e = None

    def run_e2e():
        src = SyntheticImageSource(args.batch, args.image_size, 1000) if is_worker else None
        pre = (_HostPrefetcher if dry else DevicePrefetcher)(src, device) if is_worker else None
        # the loss of every step is copied device->host (pinned, async) and READ one step later, when its copy event has
        # completed: the host never runs more than one step ahead, and it never drains the GPU queue either (a hard
        # synchronize per step leaves the GPU idle while the host enqueues the next step's first kernels)
        host_loss = [pin(torch.zeros(1, dtype=torch.float32)) for _ in range(2)]
        loss_ev = [Event(), Event()]
        for _ in range(3):
            if is_worker:
                xb, yb = pre.next()
                run_step(xb, yb)
            else:
                run_step()
        sync_all()
        e_start, e_end = Event(enable_timing=True), Event(enable_timing=True)
        last, losses_read = 0.0, 0
        e_sampler = ClockSampler(gpu_index=local_rank) if rank == 0 else None      # clocks / throttle reasons of THIS loop
        if e_sampler:
            e_sampler.__enter__()
        e_start.record()
        for i in range(args.steps):
            if is_worker:
                xb, yb = pre.next()                      # H2D of this step's inputs (pinned -> device) + normalise
                loss = run_step(xb, yb)
                k = i & 1
                host_loss[k].copy_(loss.detach().float().reshape(1), non_blocking=True)   # D2H of this step's loss
                loss_ev[k].record()
                if i > 0:                                # read the previous step's loss (its copy has finished)
                    loss_ev[k ^ 1].synchronize()
                    last = float(host_loss[k ^ 1][0])
                    losses_read += 1
            else:
                run_step()
        if is_worker:                                    # ... and the last step's, still inside the timed region
            k = (args.steps - 1) & 1
            loss_ev[k].synchronize()
            last = float(host_loss[k][0])
            losses_read += 1
            assert losses_read == args.steps
        e_end.record()
        sync_all()
        if e_sampler:
            e_sampler.__exit__(None, None, None)
        e_ms = e_start.elapsed_time(e_end)
        if os.environ.get("DLCFN_BENCH_DIAG") == "1" and is_worker:    # diagnostics only: is the e2e gap time-dependent?
            for tag, (xa, ya) in (("device input after e2e", (x, y)), ("last prefetched batch, no copies in flight", (xb, yb))):
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                d0.record()
                for _ in range(args.steps):
                    trainer.step(xa, ya)
                d1.record()
                torch.cuda.synchronize()
                print(f"[diag] {tag}: {d0.elapsed_time(d1) / args.steps:.3f} ms/step", file=sys.stderr, flush=True)
        if last != last or last in (float("inf"), float("-inf")):
            print(f"[bench] WARNING: end-to-end loss is {last}", file=sys.stderr, flush=True)
            last = None
        return {"ms": e_ms, "loss": last, "h2d": src.bytes_per_batch if is_worker else 0,
                "clocks": e_sampler.summary() if e_sampler else {}}

    if not args.no_e2e:
        if world == 1:
            try:
                e2e = run_e2e()
            except Exception as e:      # single process: the device-timed headline above is still reported
                print(f"[bench] end-to-end loop failed ({type(e).__name__}: {e}); no e2e key", file=sys.stderr, flush=True)
        else:                           # several ranks: an exception must end the job (a rank that skipped the loop would leave
            e2e = run_e2e()             # its peers waiting inside their steps)

    # ---------------------------------------------------------------- reduce over ranks
    t = torch.tensor([ms, e2e["ms"] if e2e else 0.0, float(launches)], device=device, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    else:
        e_ms = float(t[1])
    stock = None
    if args.stock and (world == 1 or args.stock >= 2) and not dry and not is_bert and args.impl == "ours" and servers is None:
        n_dev_batches = len(dev_batches)
        dev_batches = [None] * n_dev_batches            # release the device-resident inputs before the child allocates its own
        torch.cuda.empty_cache()
        sync_all()
        stock = run_stock_child(args, rank)
    if rank == 0:
        global_batch = args.batch * n_workers
        value = global_batch * args.steps / (ms / 1e3)
        clocks = sampler.summary() if sampler else {}
        par = {"local": "single", "allreduce": f"dp{world}-fused-allreduce", "ps_sync": f"ps-sync-{len(trainer.servers)}s{n_workers}w",
               "nccl": f"dp{world}-nccl-allreduce", "nccl_ps": f"dp{world}-nccl-reduce-bcast", "gloo": f"dp{world}-gloo-allreduce",
               "hier": f"dp{world}-hier-{getattr(trainer, 'num_nodes', 1)}x{trainer.node_size}"}[trainer.mode]
        if is_bert:
            rec = {"metric": f"{args.model}_pretrain_sequences_per_sec", "value": round(value, 2), "unit": "sequences/sec",
                   "tokens_per_sec": round(value * args.seq_len, 1),
                   "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
                   "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (dry run)" if dry else "bf16",
                   "data": "synthetic (random token ids, random-init weights)", "impl": "ours" if args.impl == "ours" else "nccl-baseline",
                   "config": {"model": args.model, "global_batch": global_batch, "per_gpu_batch": args.batch,
                              "seq_len": args.seq_len, "parallelism": par, "optimizer": "adam multi-precision (fp32 master)",
                              "bucket_mb": args.bucket_mb, "l2": "activations larger than L2",
                              "baseline": "none published (the reference has no transformer workload)"},
                   "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", [])},
                   "gpu_launches": launches, "last_loss": timed_loss}
            if dry:
                rec["dry_cpu"] = "host dry run of bench.py with tiny shapes - NOT a measurement"
            print(json.dumps(rec), flush=True)
            rec = None
        else:
          rec = {"metric": "resnet50_train_images_per_sec", "value": round(value, 2), "unit": "images/sec",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": round(ms / args.steps, 3), "ms_per_step_min": round(step_ms[0], 3),
               "ms_per_step_median": round(step_ms[len(step_ms) // 2], 3), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": round(value / (BASELINE_IMG_S_PER_GPU * world), 3), "dtype": "fp32 (dry run)" if dry else "bf16",
               "data": "synthetic (random uint8 images normalised on the device, random labels, random-init weights)",
               "impl": "ours" if args.impl == "ours" else "nccl-baseline",
               "config": {"model": f"resnet-{args.num_layers}-v2 (reference symbols/resnet.py)",
                          "global_batch": global_batch, "per_gpu_batch": args.batch, "image": args.image_size,
                          "seq_len": None, "parallelism": par, "optimizer": "sgd-momentum multi-precision (fp32 master)",
                          "bucket_mb": args.bucket_mb, "l2": "inputs+activations larger than L2 (no flush needed)",
                          "conv1x1": conv_ops.conv1x1_backend(), "convkxk": conv_ops.convkxk_backend(),
                          "cuda_graph": bool(graphed is not None and graphed.graphed),
                          "device_batches": len(dev_batches),
                          "baseline": "363.69 img/s per V100 (perf.md:219) x n_gpus"},
               "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"),
                          "reasons": clocks.get("reasons", [])},
               "gpu_launches": launches, "last_loss": timed_loss}
        if e2e and rec is not None:
            rec["e2e"] = {"value": round(global_batch * args.steps / (e_ms / 1e3), 2), "unit": "images/sec",
                          "h2d_bytes_per_step": e2e["h2d"] * n_workers if is_worker else 0,
                          "d2h_bytes_per_step": 4 * n_workers, "last_loss": e2e["loss"],
                          "readback": "loss copied D2H (pinned, async) every step and read on the host one step later",
                          "clocks": {"sm_mhz": e2e["clocks"].get("sm_mhz"), "reasons": e2e["clocks"].get("reasons", [])}}
            if not is_worker:  # rank 0 is a dedicated server: bytes come from the workers
                rec["e2e"]["h2d_bytes_per_step"] = (args.batch * args.image_size * args.image_size * 3 + args.batch * 8) * n_workers
        if rec is not None and stock is not None and stock.get("value"):
            # same box, same process tree, run back to back: stock nn.Module ResNet-50-v2 + torch.optim.SGD (+ DDP/NCCL), cuDNN
            # convolutions and BatchNorm, no code of this repository (baseline/torch_stock.py)
            rec["vs_stock"] = round(rec["value"] / stock["value"], 4)
            rec["stock"] = {"value": stock["value"], "ms_per_step": stock.get("ms_per_step"), "impl": "torch-stock",
                            "clocks": stock.get("clocks")}
            if "e2e" in rec and stock.get("e2e", {}).get("value"):
                rec["e2e"]["vs_stock"] = round(rec["e2e"]["value"] / stock["e2e"]["value"], 4)
                rec["stock"]["e2e_value"] = stock["e2e"]["value"]
        if rec is not None:
            if dry:
                rec["dry_cpu"] = "host dry run of bench.py with tiny shapes - NOT a measurement"
            print(json.dumps(rec), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
