"""Host-side Objective bookkeeping (theseus/core/objective.py:210-470): add / queries / erase, no GPU."""
import pytest
import torch

import theseus_b200 as th


def _objective():
    d = torch.float64
    a, b, c = (th.SE3(name=n, dtype=d) for n in "abc")
    z1, z2 = th.SE3(name="z1", dtype=d), th.SE3(name="z2", dtype=d)
    w = th.ScaleCostWeight(th.Variable(torch.ones(1, 1, dtype=d), name="w"))
    obj = th.Objective(dtype=d)
    obj.add(th.Between(a, b, z1, w, name="ab"))
    obj.add(th.Between(b, c, z2, w, name="bc"))
    obj.add(th.Difference(a, th.SE3(name="t", dtype=d), th.ScaleCostWeight(th.Variable(torch.ones(1, 1, dtype=d), name="wp")), name="prior"))
    return obj


def test_queries_and_sizes():
    obj = _objective()
    assert obj.size() == (3, 3, 5) and obj.dim() == 18 and obj.size_variables() == 3
    assert obj.has_cost_function("ab") and not obj.has_cost_function("zz") and obj.get_cost_function("bc").name == "bc"
    assert obj.has_optim_var("b") and not obj.has_optim_var("z1") and obj.has_aux_var("z1") and obj.has_aux_var("w")
    assert [cf.name for cf in obj.get_functions_connected_to_optim_var("b")] == ["ab", "bc"]
    assert [cf.name for cf in obj.get_functions_connected_to_aux_var("w")] == ["ab", "bc"]
    with pytest.raises(ValueError):
        obj.get_functions_connected_to_optim_var("nope")


def test_erase_removes_orphans_and_invalidates_structure():
    obj = _objective()
    v0 = obj._structure_version
    obj.erase("bc")
    assert obj.size() == (2, 2, 4) and not obj.has_optim_var("c") and not obj.has_aux_var("z2") and obj.has_aux_var("w")
    assert obj._structure_version > v0
    obj.erase("prior")
    assert list(obj.cost_functions) == ["ab"] and not obj.has_aux_var("wp") and not obj.has_aux_var("t")
    with pytest.warns(UserWarning):
        obj.erase("prior")


def test_duplicate_names_are_rejected_like_the_reference():
    obj = _objective()
    d = torch.float64
    with pytest.raises(ValueError, match="same name"):
        obj.add(th.Difference(th.SE3(name="a", dtype=d), th.SE3(name="t2", dtype=d), th.ScaleCostWeight(th.Variable(torch.ones(1, 1, dtype=d), name="w3")),
                              name="other"))
    with pytest.raises(ValueError, match="same name"):
        obj.add(th.Difference(obj.get_optim_var("a"), th.SE3(name="t3", dtype=d), th.ScaleCostWeight(th.Variable(torch.ones(1, 1, dtype=d), name="w4")),
                              name="prior"))


def test_random_generators_give_valid_group_elements():
    g = torch.Generator().manual_seed(0)
    for fn in (th.rand_se3, th.randn_se3):
        T = fn(5, generator=g, dtype=torch.float64).tensor
        assert T.shape == (5, 3, 4)
        R = T[:, :, :3]
        assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand(5, 3, 3), atol=1e-12)
        assert torch.allclose(torch.linalg.det(R), torch.ones(5, dtype=torch.float64), atol=1e-12)
    R = th.rand_so3(4, generator=g, dtype=torch.float64).tensor
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand(4, 3, 3), atol=1e-12)
    X = th.rand_se2(6, generator=g, dtype=torch.float64).tensor
    assert X.shape == (6, 4) and torch.allclose(X[:, 2] ** 2 + X[:, 3] ** 2, torch.ones(6, dtype=torch.float64), atol=1e-14)
    assert th.rand_vector(3, 7, generator=g).tensor.shape == (3, 7) and th.randn_point3(2, generator=g).tensor.shape == (2, 3)
    assert th.rand_point2(2, generator=g).tensor.shape == (2, 2)
