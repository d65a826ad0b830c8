"""`parse-log` - turn a training log into a table (reference MX/tools/parse_log.py): one row per epoch with the train /
validation metrics and the epoch time, from the lines the fit loops print
(`Epoch[3] Train-accuracy=0.91`, `Epoch[3] Validation-accuracy=0.88`, `Epoch[3] Time cost=12.3`), plus the mean of the
Speedometer lines (`Epoch[3] Batch [40] Speed: 1234.5 samples/sec`).

    python -m deeplearning_cfn_b200 parse-log train.log [--format markdown|none] [--metric-names accuracy top_k_accuracy_5]
"""
from __future__ import annotations

import argparse
import re
import sys

_METRIC = re.compile(r"Epoch\[(\d+)\]\s+(Train|Validation)-([\w\-.]+)=([-+\w.]+)")
_TIME = re.compile(r"Epoch\[(\d+)\]\s+Time cost=([\d.]+)")
_SPEED = re.compile(r"Epoch\[(\d+)\]\s+Batch\s*\[\d+\]\s+Speed:\s*([\d.]+)\s+samples/sec")


def parse(lines, metric_names=None) -> tuple[list[str], list[list]]:
    """-> (column names, rows); a row is [epoch, train-<m>..., val-<m>..., time, speed]."""
    data: dict[int, dict] = {}
    seen: list[str] = []
    for line in lines:
        m = _METRIC.search(line)
        if m:
            ep, kind, name, val = int(m.group(1)), m.group(2).lower()[:5], m.group(3), m.group(4)
            try:
                v = float(val)
            except ValueError:
                continue
            if metric_names and name not in metric_names:
                continue
            if name not in seen:
                seen.append(name)
            data.setdefault(ep, {})[f"{'train' if kind == 'train' else 'val'}-{name}"] = v
            continue
        m = _TIME.search(line)
        if m:
            data.setdefault(int(m.group(1)), {})["time"] = float(m.group(2))
            continue
        m = _SPEED.search(line)
        if m:
            data.setdefault(int(m.group(1)), {}).setdefault("_speeds", []).append(float(m.group(2)))
    cols = ["epoch"] + [f"train-{n}" for n in seen] + [f"val-{n}" for n in seen] + ["time", "speed"]
    rows = []
    for ep in sorted(data):
        d = data[ep]
        sp = d.get("_speeds")
        d["speed"] = sum(sp) / len(sp) if sp else None
        rows.append([ep] + [d.get(c) for c in cols[1:]])
    return cols, rows


def to_markdown(cols, rows) -> str:
    def fmt(v):
        if v is None:
            return "-"
        return str(v) if isinstance(v, int) else f"{v:.6f}".rstrip("0").rstrip(".") if abs(v) < 1e6 else f"{v:.3e}"
    out = ["| " + " | ".join(cols) + " |", "| " + " | ".join("---" for _ in cols) + " |"]
    out += ["| " + " | ".join(fmt(v) for v in r) + " |" for r in rows]
    return "\n".join(out)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Parse a training log into a table")
    ap.add_argument("logfile", nargs=1)
    ap.add_argument("--format", default="markdown", choices=["markdown", "none"])
    ap.add_argument("--metric-names", nargs="+", default=None, help="only these metrics (default: every metric in the log)")
    a = ap.parse_args(argv)
    with open(a.logfile[0]) as f:
        cols, rows = parse(f, a.metric_names)
    if a.format == "markdown":
        print(to_markdown(cols, rows))
    else:
        print("\t".join(cols))
        for r in rows:
            print("\t".join("" if v is None else str(v) for v in r))
    return 0


if __name__ == "__main__":
    sys.exit(main())
