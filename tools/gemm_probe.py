"""Runs each tcgen05 GEMM variant in its own subprocess (a hung kernel only costs that variant).
Usage: python tools/gemm_probe.py [--perf]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from deeplearning_cfn_b200.ops import native
M,N,K,a_mn,b_mn,bn,perf = %s
g = torch.Generator().manual_seed(0)
a = (torch.randn(M,K,generator=g)*0.5).cuda().bfloat16(); b = (torch.randn(N,K,generator=g)*0.5).cuda().bfloat16()
A = a.t().contiguous() if a_mn else a; B = b.t().contiguous() if b_mn else b
out = native.gemm_bf16(A,B,a_mn=bool(a_mn),b_mn=bool(b_mn),block_n=bn); torch.cuda.synchronize()
want = a.float() @ b.float().t()
err = ((out.float()-want).abs().max()/(want.abs().max()+1e-6)).item()
res = {"err": err}
if perf:
    for _ in range(3): native.gemm_bf16(A,B,a_mn=bool(a_mn),b_mn=bool(b_mn),block_n=bn,out=out)
    s,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): native.gemm_bf16(A,B,a_mn=bool(a_mn),b_mn=bool(b_mn),block_n=bn,out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)/20; res["ms"]=ms; res["tflops"]=2*M*N*K/ms/1e9
    s.record()
    for _ in range(20): torch.matmul(a, b.t())
    e.record(); torch.cuda.synchronize(); ms2 = s.elapsed_time(e)/20; res["cublas_tflops"]=2*M*N*K/ms2/1e9
print("RESULT"+json.dumps(res))
'''


def main():
    perf = "--perf" in sys.argv
    cases = []
    for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        for bn in (128, 256):
            cases.append((256, 512, 256, a_mn, b_mn, bn, 0))
    cases += [(300, 1000, 2048, 0, 0, 128, 0), (1000, 264, 72, 0, 0, 128, 0)]
    if perf:
        for (M, N, K) in ((8192, 8192, 8192), (4096, 4096, 1024), (4096, 1024, 4096), (16384, 1024, 1024)):
            for bn in (128, 256):
                cases.append((M, N, K, 0, 0, bn, 1))
        cases += [(8192, 8192, 8192, 0, 1, 256, 1), (8192, 8192, 8192, 1, 1, 256, 1)]
    out = []
    for c in cases:
        code = CHILD % (ROOT, repr(c))
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=90)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            if line:
                res = json.loads(line[0][6:])
                res["status"] = "ok" if res["err"] < 2e-2 else "WRONG"
            else:
                res = {"status": "CRASH", "tail": (r.stdout + r.stderr)[-400:]}
        except subprocess.TimeoutExpired:
            res = {"status": "HANG"}
        res["case"] = c
        print(json.dumps(res), flush=True)
        out.append(res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
