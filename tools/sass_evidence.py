#!/usr/bin/env python
"""SASS evidence for the tensor-core kernels of libasep.so: per kernel the instruction count, the counts of the
Blackwell-native opcodes and every line that carries one.   python tools/sass_evidence.py > profiles/r02_sass_tensor_kernels.txt"""
import collections
import re
import subprocess
import sys

LIB = "audiosourcesep_b200/libasep.so"
KERNELS = ("k_nn_tc4", "k_nn_tcx", "k_conv_tc", "k_conv_tc_sw", "k_wgrad_tc", "k_conv_wgrad_tc")
OPS = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UTCATOMSWS",
       "UTMACMDFLUSH", "UTCCP")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print("SASS evidence for the tensor-core kernels of libasep.so (cuobjdump -sass audiosourcesep_b200/libasep.so, sm_100a; round 2).")
    print("Per kernel: instruction count, the Blackwell-native opcodes (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,")
    print("UTMALDG/UTMASTG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and every line that carries a tensor / TMA opcode.\n")
    name, lines = None, []

    def flush():
        if name is None or not any(k + "I" in name or name.endswith(k + "ENS") or ("_" + k) in name or k in name for k in KERNELS):
            return
        if not any(re.search(r"\d" + re.escape(k) + r"(I|E)", name) for k in KERNELS):
            return
        cnt = collections.Counter()
        keep = []
        for ln in lines:
            m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if not m:
                continue
            cnt["_n"] += 1
            op = m.group(1).split(".")[0]
            if op in OPS:
                cnt[op] += 1
                if op not in ("SYNCS",):
                    keep.append(ln.rstrip().split("/* 0x")[0].rstrip())
        print("== " + name)
        print(f"   {cnt['_n']} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(cnt.items()) if k != "_n"))
        for ln in keep[:60]:
            print("  " + ln)
        if len(keep) > 60:
            print(f"   ... {len(keep) - 60} more tensor / TMA lines")
        print()

    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            name, lines = m.group(1), []
        elif name is not None:
            lines.append(ln)
    flush()


if __name__ == "__main__":
    main()
