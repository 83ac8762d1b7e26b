"""Builds tools/x87_check.cpp around the SlabX87 section of sla_b200/csrc/slab_common.cuh and runs it:
4 million random accumulations compared bit for bit with native long double (x86-64 only)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "sla_b200", "csrc", "slab_common.cuh")).read()
a = src.index("__host__ __device__ __forceinline__ int slab_clz64")
b = src.index("/* ---------------- CRC-16/IBM")
with tempfile.TemporaryDirectory() as tmp:
    open(os.path.join(tmp, "x87_part.h"), "w").write(src[a:b])
    exe = os.path.join(tmp, "x87_check")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-I", tmp, "-o", exe, os.path.join(ROOT, "tools", "x87_check.cpp")], check=True)
    sys.exit(subprocess.run([exe]).returncode)
