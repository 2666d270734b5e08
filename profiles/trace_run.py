import os, sys, tarfile, tempfile, time
ROOT='/root/repo'
sys.path.insert(0, ROOT)
from catt_b200.host import Catt
from catt_b200.refg import chain_config
n=int(sys.argv[1])
d = tempfile.mkdtemp(prefix="catt_tr_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
cc = chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource"))
ctx = Catt(chain_cfg=cc, device=0, only_cdr3=True)
bufs = ctx.synth_reads(n, 150)
for i in range(3):
    ctx.mainflow_buffers(n, *bufs).close()
os.environ["CATT_TRACE"]="1"; os.environ["CATT_TRACE_CHUNKS"]="1"
t0=time.perf_counter()
r=ctx.mainflow_buffers(n, *bufs)
print("wall ms", (time.perf_counter()-t0)*1e3)
r.close()
