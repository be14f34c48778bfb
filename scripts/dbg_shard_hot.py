"""debug aid: shard a multi-component graph on ONE device and run the hot path on every share"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from povu_b200 import api, synth
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pv = api.Povu(device=0) if not os.environ.get('PV_EMU_LIB') else api.Povu(lib_path=os.environ['PV_EMU_LIB'])
simple = [synth.template_from(synth.snp_indel_chain(6, n_paths=6, seed=100 + k)) for k in range(18)] + synth.cfg3_templates(seed=7, n_nested=2, tower_depth=0)
W = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
     114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
g = synth.tiled(simple, n_inst, 24, seed=11, n_components=25, comp_weights=W)
for rank in range(world):
    s = api.Session.from_graph(pv, g)
    print('shard', s.shard(rank, world), s.sizes(), flush=True)
    ga = s.graph_arrays(labels=False)
    own = [n for k, n in enumerate(ga["path_names"]) if n.startswith("S001#1#") and ga["step_off"][k + 1] > ga["step_off"][k]]
    print(own, flush=True)
    print(s.hot_path([own[0]], keep_resident=True), s.lib.pvgpu_session_record_count(s.h), flush=True)
    for p in own[1:3]:
        print(p, s.call([p]), flush=True)
    s.close()
