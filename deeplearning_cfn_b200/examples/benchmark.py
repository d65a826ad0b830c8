"""Scaling sweep driver.

Reference: `MX/example/image-classification/benchmark.py` - for each network run GPUs 1,2,4..G on one node with
kv-store=device, then nodes 2,4..N with dist_sync_device through `launch.py -n N -s 2N -H hosts`
(:186-222, 240-258), scrape the `Speed:` lines (:156-166), write `benchmark/<net>.csv` and a speed-up chart
(:224-238); `--networks mode:name:batch:imgsize...`, `--worker_file`, `--worker_count`, `--gpu_count` (:97-141).
Here the whole sweep stays inside one NVSwitch box: N = 1, 2, 4, ... `--gpu_count` ranks of
`examples.train_imagenet --benchmark 1`, for each `--kv-stores` entry, device-side throughput from the
Speedometer lines; the chart is a dependency-free SVG.
"""
from __future__ import annotations

import argparse
import csv
import os
import re
import subprocess
import sys

SPEED_RE = re.compile(r"Speed: ([0-9.]+) samples/sec")


def parse(argv=None):
    p = argparse.ArgumentParser(description="Run a scaling benchmark over networks / GPU counts / kv-stores")
    p.add_argument("--worker_file", type=str, default=os.environ.get("DEEPLEARNING_WORKERS_PATH", "/opt/deeplearning/workers"))
    p.add_argument("--worker_count", type=int, default=int(os.environ.get("DEEPLEARNING_WORKERS_COUNT", "1")))
    p.add_argument("--gpu_count", type=int, default=int(os.environ.get("DEEPLEARNING_WORKER_GPU_COUNT", "0")) or 8)
    p.add_argument("--networks", type=str, nargs="+", default=["native:resnet:256:224"],
                   help="mode:network[-layers]:batch_per_gpu:image_size")
    p.add_argument("--kv-stores", type=str, nargs="+", default=["allreduce", "dist_sync_device", "nccl"])
    p.add_argument("--num-batches", type=int, default=30)
    p.add_argument("--out-dir", type=str, default="benchmark")
    p.add_argument("--dry-run", action="store_true", help="print the commands without running them")
    return p.parse_args(argv)


def gpu_steps(max_gpus: int) -> list[int]:
    n, out = 1, []
    while n <= max_gpus:
        out.append(n)
        n *= 2
    return out


def command(net: str, layers: int, batch: int, size: int, kv: str, n: int, num_batches: int, cpu: bool = False) -> list[str]:
    train = [sys.executable, "-m", "deeplearning_cfn_b200.examples.train_imagenet", "--benchmark", "1", "--network", net,
             "--num-layers", str(layers), "--batch-size", str(batch), "--image-shape", f"3,{size},{size}", "--kv-store", kv,
             "--num-epochs", "1", "--num-batches", str(num_batches), "--disp-batches", str(max(1, min(5, num_batches // 2)))]
    if cpu:
        train += ["--gpus", ""]
    if n == 1:
        return train
    return [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", str(n)] + train


def scrape(log: str, skip: int = 1) -> float:
    """Mean of the Speed lines, skipping the first (warm-up / autotune) ones; per-rank speeds are summed by the caller."""
    v = [float(m.group(1)) for m in SPEED_RE.finditer(log)]
    v = v[skip:] if len(v) > skip else v
    return sum(v) / len(v) if v else 0.0


def svg_chart(series: dict, xs: list[int], path: str, title: str) -> None:
    w, h, m = 640, 400, 50
    ymax = max([max(v) for v in series.values()] + [max(xs)]) * 1.1
    X = lambda x: m + (w - 2 * m) * (xs.index(x) / max(1, len(xs) - 1))   # noqa: E731
    Y = lambda y: h - m - (h - 2 * m) * (y / ymax)                        # noqa: E731
    col = ["#1f77b4", "#d62728", "#2ca02c", "#9467bd", "#ff7f0e"]
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{w}" height="{h}"><rect width="100%" height="100%" fill="white"/>',
           f'<text x="{w / 2}" y="20" text-anchor="middle" font-size="14">{title}</text>',
           f'<line x1="{m}" y1="{h - m}" x2="{w - m}" y2="{h - m}" stroke="black"/><line x1="{m}" y1="{m}" x2="{m}" y2="{h - m}" stroke="black"/>']
    for x in xs:
        out.append(f'<text x="{X(x)}" y="{h - m + 16}" text-anchor="middle" font-size="11">{x}</text>')
    ideal = " ".join(f"{X(x)},{Y(x)}" for x in xs)
    out.append(f'<polyline points="{ideal}" fill="none" stroke="#999" stroke-dasharray="4"/>')
    for i, (name, vals) in enumerate(series.items()):
        pts = " ".join(f"{X(x)},{Y(v)}" for x, v in zip(xs, vals))
        out.append(f'<polyline points="{pts}" fill="none" stroke="{col[i % 5]}" stroke-width="2"/>')
        out.append(f'<text x="{w - m - 150}" y="{m + 16 * i}" font-size="11" fill="{col[i % 5]}">{name}</text>')
    out.append("</svg>")
    open(path, "w").write("\n".join(out))


def main(argv=None) -> int:
    a = parse(argv)
    os.makedirs(a.out_dir, exist_ok=True)
    cpu = a.gpu_count <= 0            # a box without GPUs: sweep the number of CPU ranks per worker instead
    xs = gpu_steps(max(1, a.gpu_count) * max(1, a.worker_count))
    for spec in a.networks:
        try:
            _, net, batch, size = spec.split(":")
            layers = 50
            if "-" in net:
                net, l = net.split("-")
                layers = int(l)
            int(batch), int(size)
        except ValueError:
            raise SystemExit(f"benchmark: --networks takes mode:network[-layers]:batch_per_gpu:image_size (e.g. native:resnet-50:256:224), "
                             f"got {spec!r}")
        rows, speedup = [], {}
        for kv in a.kv_stores:
            speeds = []
            for n in xs:
                cmd = command(net, layers, int(batch), int(size), kv if n > 1 else "device", n, a.num_batches, cpu)
                if a.dry_run:
                    print(" ".join(cmd))
                    speeds.append(0.0)
                    continue
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    print(f"{net}-{layers} kv={kv} ranks={n}: run failed (exit {r.returncode})\n" + (r.stdout + r.stderr)[-800:],
                          file=sys.stderr, flush=True)
                per_rank = scrape(r.stdout + r.stderr)
                total = per_rank * n
                speeds.append(total)
                rows.append({"network": net, "layers": layers, "kv_store": kv, "gpus": n, "images_per_sec": round(total, 2)})
                print(f"{net}-{layers} kv={kv} gpus={n}: {total:.1f} img/s", flush=True)
            if speeds and speeds[0] > 0:
                speedup[kv] = [s / speeds[0] for s in speeds]
        if rows:
            with open(os.path.join(a.out_dir, f"{net}.csv"), "w", newline="") as f:
                wri = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
                wri.writeheader()
                wri.writerows(rows)
            if speedup:
                svg_chart(speedup, xs, os.path.join(a.out_dir, f"{net}_speedup.svg"), f"{net}-{layers} speed-up vs GPUs")
    return 0


if __name__ == "__main__":
    sys.exit(main())
