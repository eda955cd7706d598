"""GPU debug: engine beam_v vs oracle beam_v, first differing step."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from deepxube_b200 import _lib as L
import golden_utils as gu
from oracle.beam import OracleBeamV
from oracle.domains_np import OracleDomain
from oracle.pseudo_heur import pseudo_v_fine
tag = sys.argv[1] if len(sys.argv) > 1 else "beam_cube3_20B"
c = gu.beam_case(tag)
name, dim = gu.parse_domain(str(c["cfg"][0]))
b, _ = gu.parse_pathfind(str(c["cfg"][1]))
dom = OracleDomain(name, dim)
n = c["states"].shape[0]
orc = OracleBeamV(dom, pseudo_v_fine, b)
oi = orc.add_instances(c["states"], c["goals"])
eng = L.Engine(name, dim, "beam_v", L.DX_HEUR_HOST, max_instances=n, max_pop_total=n * b, max_nodes=400000)
eng.add_instances(c["states"], c["goals"], b, 1.0, root_vals=pseudo_v_fine(c["states"], c["goals"]))
for t in range(c["itr"].shape[0]):
    orc.step()
    npend = eng.step_begin()
    ps, pg = eng.get_pending(npend)
    eng.step_end(pseudo_v_fine(ps, pg))
    for j, s in enumerate(eng.status()):
        st, g, h = eng.curr_nodes(j, 4096)
        ids = oi[j].nodes_curr
        ost, oh = oi[j].store.state[ids], oi[j].store.h[ids]
        a = sorted(map(bytes, st)); o = sorted(map(bytes, ost))
        print(f"t={t} j={j} engine n={len(st)} oracle n={len(ids)} pending={npend} same_set={a == o} gen={s.num_nodes_generated}/{oi[j].num_nodes_generated}")
        if a != o:
            print(" engine h sorted:", np.sort(h)[:25])
            print(" oracle h sorted:", np.sort(oh)[:25])
            print(" engine h == pseudo(engine states):", np.array_equal(h.astype(np.float64), pseudo_v_fine(st)))
            from collections import Counter
            ca, co = Counter(a), Counter(o)
            for k in set(ca) | set(co):
                if ca[k] != co[k]:
                    row = np.frombuffer(k, np.uint8)
                    print("  row", row[:12], "... engine x", ca[k], "oracle x", co[k], "h", pseudo_v_fine(row[None, :])[0],
                          "colour counts", np.bincount(row, minlength=6))
            sys.exit(0)
