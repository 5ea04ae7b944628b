"""dev: compare the GPU embed of fixture rows with the reference's embed (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
if __name__ == "__main__":
    import prior_ref
    from test_theta_parity import _pipeline
    from xpsi_b200 import synthetic as syn
    rows = [int(a) for a in sys.argv[1:]] or [2]
    d = np.load(os.path.join(ROOT, "tests/golden/m2_prior.npz"))
    thetas = d["thetas"][rows]
    res = prior_ref.run_reference(thetas, full=True)
    pipe, _ = _pipeline(len(rows))
    pipe.embed_spots(syn.m2_spot_batch(pipe, thetas))
    e = pipe.fetch_embed(len(rows))
    np.set_printoptions(linewidth=200, precision=6)
    for b, r in enumerate(res):
        print("row", rows[b], "theta", thetas[b])
        for m, mem in enumerate(r["members"]):
            q = 2 * b + m
            n = mem["cellArea"].shape[0]
            A, Ar = e["cellArea"][q, :n, :n], mem["cellArea"]
            full = Ar.max()
            dA = (A - Ar) / full
            bad = np.argwhere(np.abs(dA) > 1e-10)
            print(" member", m, "rings", n, "sum area rel diff", (A.sum() - Ar.sum()) / Ar.sum(), "max cell diff/full", np.abs(dA).max())
            th = thetas[b]
            st = syn.SpacetimeScalars(th[0], th[1], th[2], th[3], syn.M2_FREQUENCY)
            ex = prior_ref.exact_spot_area((st.epsilon, st.zeta, st.R, th[5 + 4 * m], th[6 + 4 * m]))
            print("   vs exact spot area: ours rel", float((A.sum() - ex) / ex), " reference rel", float((Ar.sum() - ex) / ex))
            for i, j in bad[:20]:
                print("   cell", i, j, "ours", A[i, j] / full, "ref", Ar[i, j] / full, "diff/full", dA[i, j],
                      "theta", mem["theta"][i, 0], "phi", mem["phi"][i, j])
            for k in ("deflection", "cos_alpha", "lag"):
                dd = np.abs(e[k][q, :n] - mem[k])
                i, j = np.unravel_index(np.argmax(dd), dd.shape)
                print("  ", k, "max abs diff", dd.max(), "at ring", i, "ray", j, "value", mem[k][i, j])
            dm = np.abs(e["maxDeflection"][q, :n] - mem["maxDeflection"])
            print("   maxDeflection max abs diff", dm.max(), "value", mem["maxDeflection"][np.argmax(dm)])
