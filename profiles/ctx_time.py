import os, sys, tarfile, tempfile, time
sys.path.insert(0, '/root/repo')
t0=time.perf_counter()
from catt_b200.host import Catt
from catt_b200.refg import chain_config
d = tempfile.mkdtemp()
with tarfile.open('/root/repo/tests/golden/catt_fixture.tar.gz') as tar: tar.extractall(d)
cc = chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource"))
t1=time.perf_counter()
os.environ["CATT_TRACE"]="1"
c = Catt(chain_cfg=cc, device=0, only_cdr3=True)
t2=time.perf_counter()
c2 = Catt(chain_cfg=cc, device=0, only_cdr3=True)
t3=time.perf_counter()
r = c.mainflow(os.path.join(d, "testSample.fq")); t4=time.perf_counter()
print(f"imports+fixture {t1-t0:.2f} s, first ctx (CUDA context + references) {1e3*(t2-t1):.0f} ms, second ctx {1e3*(t3-t2):.0f} ms, first call on testSample.fq (7,922 reads) {1e3*(t4-t3):.0f} ms")
