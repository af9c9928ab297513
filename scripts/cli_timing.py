"""`dp_map` (the C++ stand-in for the Go host) end to end on a FASTA file in /dev/shm: wall time incl. file parsing, index
build and PAF printing. N_READS x 10 kb reads against the 4.6 Mb reference (BASELINE config 1 shape by default)."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
n = int(os.environ.get("N_READS", 10000)); L = 10000
d = "/dev/shm/dp_cli"; os.makedirs(d, exist_ok=True)
ref = synth.reference(1, 4_600_000)
rd = synth.reads(ref, 11, n, L)
with open(d + "/ref.fasta", "wb") as f:
    f.write(b">ref\n"); f.write(ref.tobytes()); f.write(b"\n")
t = time.time()
with open(d + "/reads.fasta", "wb") as f:
    buf = bytearray()
    for i in range(n):
        buf += b">read%d\n" % i; buf += rd[i * L:(i + 1) * L].tobytes(); buf += b"\n"
        if len(buf) > (64 << 20): f.write(buf); buf = bytearray()
    f.write(buf)
print("wrote %d reads in %.1f s" % (n, time.time() - t), flush=True)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "dp_map")
for it in range(2):
    t = time.time()
    p = subprocess.run([exe, "-input", d + "/reads.fasta", "-reference", d + "/ref.fasta"], stdout=open(d + "/out.paf", "wb"),
                       stderr=subprocess.PIPE, env=dict(os.environ, DOWNPORE_STATS="1"))
    dt = time.time() - t
    print("dp_map: %.2f s wall -> %.2f Gbp/s; PAF lines %d" % (dt, n * L / dt / 1e9, sum(1 for _ in open(d + "/out.paf", "rb"))), flush=True)
    print(p.stderr.decode()[-1500:])
