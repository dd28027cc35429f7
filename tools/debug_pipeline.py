"""GPU debugging aid: determinism of the host-buffer pipeline vs a single context."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from protocols.secure_comparison_b200 import SchemeEngine
from protocols.secure_comparison_b200.keyfile import schemes_from_json
from protocols.secure_comparison_b200.pipeline import LocalPipeline

B, l = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 16
paillier, dgk = schemes_from_json(bench.key_file(2048, 2048, 16))
eng = SchemeEngine(0); eng.load_paillier(paillier); eng.load_dgk(dgk, l)
w, truth = bench.make_workload(B, l, 7, eng)

def single(e):
    z_enc = e.initiator_step1(w["x"], w["y"], w["r"], w["rho_z"], l)
    z = e.paillier_decrypt(z_enc)
    db = e.keyholder_step4ab(z, l, w["rho_db"])
    c = e.initiator_step4(db, w["r"], w["delta_a"], w["r_blind"], w["perm"], w["rho_c"], l)
    _, enc3 = e.keyholder_step4j_5(c, z, w["rho_enc3"], l)
    return {"z_enc": z_enc, "z": z, "db": db, "c": c, "enc3": enc3, "res": e.initiator_step6_7(enc3, w["r"], w["delta_a"], l)}

a = single(eng); b = single(eng)
print("single ctx repeatable:", {k: bool(np.array_equal(a[k], b[k])) for k in a})
bits = eng.paillier_decrypt(a["res"])
print("single ctx bits ok:", bool(np.array_equal(bits[:, 0].astype(np.uint8), truth)))
for nctx, chunks in ((1, 1), (1, 4), (2, 2), (2, 4)):
    pipe = LocalPipeline(paillier, dgk, l, n_contexts=nctx, chunks=chunks)
    for fp64 in (True, False):
        for e in pipe.engines: e.set_modexp_fp64(fp64)
        r1 = pipe.compare(w).copy(); r2 = pipe.compare(w).copy()
        bad = np.nonzero((r1 != a["res"]).any(axis=1))[0]
        print(f"ctx={nctx} chunks={chunks} fp64={fp64}: equal to single={len(bad) == 0} repeatable={bool(np.array_equal(r1, r2))} bad rows={bad[:8].tolist()} n_bad={len(bad)}")
    pipe.close()
