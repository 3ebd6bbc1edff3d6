"""The functor boundary (SURVEY §8b "functor kernels", north_star: per-element / per-intersection functors as batched device
kernels behind the C ABI): dgb_for_each_element / dgb_for_each_intersection with the compiled-in registry of
csrc/dgb_functors.cuh. The shipped functors are the three loops of the reference's doc/recipes/recipe-integration.cc:90-126
(config 1 of BASELINE.json); each integral comes from ONE fused sweep+reduce kernel and is compared with the same loop over
the oracle's per-entity values."""
import numpy as np
import pytest

from tests.common import best_oracle, make_cfg, oracle_world, product_grid


def test_functor_registry_lists_the_recipe_functors():
    import dune_grid_b200 as dg
    reg = dg.functors()
    assert set(reg) == set(dg.FUNCTORS.values())
    assert reg[dg.FUNCTORS["midpoint_exp_norm"]][1:] == (True, False)
    assert reg[dg.FUNCTORS["gauss5_exp_norm"]][1:] == (True, False)
    assert reg[dg.FUNCTORS["boundary_flux_x"]][1:] == (False, True)
    assert reg[dg.FUNCTORS["measure"]][1:] == (True, True)
    assert all("recipe-integration" in reg[i][0] for i in (0, 1, 2))


def recipe_references(w, d):
    """the three loops of recipe-integration.cc over the oracle's per-entity values, in the recipe's order"""
    r = w.entities(0, 0, 0)
    ref1 = 0.0
    for c, v in zip(r["center"], r["volume"]):
        ref1 += np.exp(np.sqrt(c @ c)) * v
    x1 = np.array([0.5 - 0.5 * np.sqrt(0.6), 0.5, 0.5 + 0.5 * np.sqrt(0.6)])
    w1 = np.array([5.0, 8.0, 5.0]) / 18.0
    if d == 2:
        qp = np.array([[a, b] for b in x1 for a in x1]); wq = np.array([wa * wb for wb in w1 for wa in w1])
    else:
        qp = np.array([[a, b, c] for c in x1 for b in x1 for a in x1]); wq = np.array([wa * wb * wc for wc in w1 for wb in w1 for wa in w1])
    rg = w.geometry_eval(0, 0, qp)
    ref2 = 0.0
    for e in range(rg["global"].shape[0]):
        for q in range(len(wq)):
            x = rg["global"][e, q]
            ref2 += np.exp(np.sqrt(x @ x)) * rg["detJ"][e, q] * wq[q]
    ri = w.intersections(0, 0)
    ref3 = 0.0
    for e in range(ri["neighbor"].shape[0]):
        for k in range(2 * d):
            if not ri["neighbor"][e, k]:
                ref3 += (ri["center"][e, k] @ ri["normal"][e, k]) * ri["volume"][e, k]
    return ref1, ref2, ref3, r, ri


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [make_cfg(dim=2, N=(256, 256)), make_cfg(dim=3, N=(20, 17, 13), mode=2),
                                 make_cfg(dim=2, N=(40, 30), periodic=1, upper=(2.0, 1.5))], ids=["config1", "tensor3d", "periodic2d"])
def test_recipe_integrals_from_one_fused_kernel_each(cfg):
    """midpoint rule, order-5 Gauss rule and boundary flux of recipe-integration.cc: tolerance 1e-12 relative (the device kernel
    sums in its fixed block order, the recipe's loop in iteration order); two runs give the same bits (deterministic reduction);
    per-entity values agree with the loop body evaluated on the oracle's entities"""
    import torch
    import dune_grid_b200 as dg
    d = cfg["dim"]
    o = best_oracle()
    w = oracle_world(o, cfg, 1)
    ref1, ref2, ref3, r, ri = recipe_references(w, d)
    g = product_grid(cfg, 0, 1)
    gv = g.leafGridView()
    s1, v1 = gv.forEachElement("midpoint_exp_norm", values=True)
    s2 = gv.forEachElement("gauss5_exp_norm")
    s3, v3 = gv.forEachIntersection("boundary_flux_x", values=True)
    assert abs(s1.item() - ref1) <= 1e-12 * abs(ref1)
    assert abs(s2.item() - ref2) <= 1e-12 * abs(ref2)
    assert abs(s3.item() - ref3) <= 1e-12 * max(1.0, abs(ref3))
    assert abs(s1.item() - s2.item()) <= 1e-3 * abs(s2.item())                        # the two rules agree to discretisation error
    body1 = np.exp(np.linalg.norm(r["center"], axis=1)) * r["volume"]
    assert np.abs(v1.cpu().numpy() - body1).max() <= 1e-14 * np.abs(body1).max()      # exp() of the device vs libm: <= 2 ulp
    body3 = np.where(ri["neighbor"] == 0, np.einsum("efd,efd->ef", ri["center"], ri["normal"]) * ri["volume"], 0.0)
    assert np.array_equal(v3.cpu().numpy().T, body3)                                  # no transcendental: bit for bit
    for name, faces in (("midpoint_exp_norm", False), ("gauss5_exp_norm", False), ("boundary_flux_x", True)):
        f = gv.forEachIntersection if faces else gv.forEachElement
        a, b = f(name), f(name)
        assert torch.equal(a, b)                                                      # reproducible summation order
    # measure functor: |domain| and total face area with a scaling argument; interior partition of a one-rank grid = all cells
    vol = gv.forEachElement("measure", [2.0]).item()
    assert abs(vol - 2.0 * float(np.sum(r["volume"]))) <= 1e-12 * vol
    area = gv.forEachIntersection("measure", [1.0], partition=dg.Interior_Partition).item()
    w_int = w.intersections(0, 0, dg.Interior_Partition)
    assert abs(area - float(w_int["volume"].sum())) <= 1e-12 * area
    assert gv.forEachElement("measure", [1.0], partition=dg.Ghost_Partition).item() == 0.0     # empty range
    w.close()


@pytest.mark.gpu
def test_functor_errors():
    import dune_grid_b200 as dg
    from dune_grid_b200 import capi
    g = product_grid(make_cfg(dim=2, N=(8, 8)), 0, 1)
    gv = g.leafGridView()
    with pytest.raises(capi.DgbError) as e:
        gv.forEachElement(77)
    assert e.value.code == capi.EARG
    with pytest.raises(capi.DgbError) as e:
        gv.forEachIntersection("midpoint_exp_norm")           # has no intersection body
    assert e.value.code == capi.EARG
    with pytest.raises(capi.DgbError) as e:
        gv.forEachElement("boundary_flux_x")                  # has no element body
    assert e.value.code == capi.EARG
