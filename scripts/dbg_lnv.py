import sys, faulthandler
faulthandler.enable()
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from povu_b200 import api
import helpers
import os
pv = api.Povu(lib_path=os.environ["PV_LIB"]) if os.environ.get("PV_LIB") else api.Povu(device=0)
for name in sys.argv[1:]:
    s = api.Session(pv, helpers.fixture_path(name))
    print("open", flush=True)
    s.decompose(); print("decompose", flush=True)
    s.call(["HG1"]); print("call", flush=True)
    t = s.vcf_text(); print("text", len(t), flush=True)
    del s
    print("closed", flush=True)
