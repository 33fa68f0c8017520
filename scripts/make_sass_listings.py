"""Regenerate profiles/sass/*.sass (+ opcode summary) from the built native library."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "mpi4jax_b200", "_native", "libb2mpi.so")
OUT = os.path.join(REPO, "profiles", "sass")
WANT = {
    "b2_k_halo": "halo", "swe_k1": "swe_k1_fluxes", "swe_k2": "swe_k2_tendencies", "swe_k3": "swe_k3",
    "swe_k4": "swe_k4", "swe_k5": "swe_k5", "b2_k_barrier": "barrier", "b2_k_move": "move",
    "b2_k_p2p": "p2p", "b2_k_allreduce_nvlsILi0": "allreduce_nvls_f32",
    "b2_k_allreduce_nvlsILi3": "allreduce_nvls_bf16", "b2_k_allreduce_llIfLi0": "allreduce_ll_f32_sum",
    "b2_k_reduce_chunkedIfLi0": "reduce_chunked_f32_sum",
    "b2_k_reduce_chunkedI13__nv_bfloat16Li0": "reduce_chunked_bf16_sum",
    "b2_k_gemm": "gemm_allreduce_tcgen05",
}
HEADER = """# SASS listings (cuobjdump -sass libb2mpi.so, sm_100a)

One file per kernel family (typed reduce kernels: the f32/bf16 SUM instances); instruction
encodings stripped.  What to look for: peer/multicast traffic in the same kernel as the
arithmetic -- `LDG.E.128.STRONG.SYS` (peer pulls), `STG.E.128` on peer-mapped pointers
(pushes), `LDGMC.E.ADD.F32x4` / `LDGMC.E.HPADD.BF16x8` (= `multimem.ld_reduce`, in-switch
reduction), `REDG.E.ADD.STRONG.SYS` (arrival counters), `MEMBAR.*.SYS` (release fences).
The collective and stencil kernels contain no UTC*MMA / UTMALDG: they are bandwidth- or
latency-bound data movers / streaming stencils (~1 FLOP per 2-4 bytes) with no GEMM-shaped work.

| file | mangled name | instructions | memory / sync opcodes |
|---|---|---|---|
"""


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    os.makedirs(OUT, exist_ok=True)
    rows = []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        for key, out in WANT.items():
            if key in name:
                body = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/", "", f)
                ops = collections.Counter(
                    re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", body, re.M))
                keep = {k: v for k, v in ops.items()
                        if re.match(r"(LDG|STG|LDGMC|RED|ATOM|MEMBAR|CCTL|BAR|LDS|STS|UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS)", k)}
                with open(os.path.join(OUT, out + ".sass"), "w") as fh:
                    fh.write("Function : " + body)
                rows.append((out, name, sum(ops.values()), keep))
    with open(os.path.join(OUT, "README.md"), "w") as fh:
        fh.write(HEADER)
        for out, name, n, keep in sorted(rows):
            fh.write(f"| {out}.sass | `{name[:70]}` | {n} | "
                     f"{', '.join(f'{k} x{v}' for k, v in sorted(keep.items()))} |\n")
    print(f"wrote {len(rows)} listings to {OUT}")


if __name__ == "__main__":
    sys.exit(main())
