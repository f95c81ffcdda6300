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


def _pair_objective(num_variables, rng):
    vs = [th.Vector(1, name=f"var{i}") for i in range(num_variables)]
    pairs = [(a, b) for i, a in enumerate(vs) for b in vs[i + 1:]]
    rng.shuffle(pairs)
    objective, expected = th.Objective(), []
    w = th.ScaleCostWeight(1.0)
    for k, (a, b) in enumerate(pairs):
        objective.add(th.AutoDiffCostFunction([a, b], lambda optim_vars, aux_vars: optim_vars[0].tensor - optim_vars[1].tensor, 1, cost_weight=w,
                                              name=f"cf{k}"))
        for v in (a, b):
            if v not in expected:
                expected.append(v)
    return objective, vs, expected


def test_variable_ordering_default_append_remove_iterate():
    """The reference's own checks (tests/theseus_tests/optimizer/test_variable_ordering.py): default order = first appearance; append /
    remove / extend / complete / iteration of a hand-built order."""
    import random
    rng = random.Random(0)
    for n in range(2, 9):
        objective, vs, expected = _pair_objective(n, rng)
        order = th.VariableOrdering(objective)
        assert [order.index_of(v.name) for v in expected] == list(range(n)) and order.complete
    objective, vs, _ = _pair_objective(8, rng)
    for _ in range(10):
        rng.shuffle(vs)
        order = th.VariableOrdering(objective, default_order=False)
        assert not order.complete
        order.extend(vs[:3])
        for v in vs[3:]:
            order.append(v)
        assert order.complete and [order.index_of(v.name) for v in vs] == list(range(8)) and [v for v in order] == vs and order[2] is vs[2]
        with pytest.raises(ValueError):
            order.append(vs[0])
        rng.shuffle(vs)
        for v in vs:
            order.remove(v)
            assert not order.complete and v not in order._var_order and v.name not in order._var_name_to_index
        with pytest.raises(ValueError):
            order.append(th.Vector(1, name="stranger"))


def test_custom_variable_ordering_is_recorded_for_the_engine():
    """A Linearization with a custom VariableOrdering records the column order on the objective (the engine is compiled for it on first
    use, on the GPU); the default order records nothing; an incomplete order is refused like in the reference (linearization.py:29-30)."""
    import random
    from theseus_b200.optimizer import Linearization
    objective, vs, expected = _pair_objective(5, random.Random(1))
    lin = Linearization(objective)
    assert objective._engine_ordering is None and lin.var_start_cols == [0, 1, 2, 3, 4]
    order = th.VariableOrdering(objective, default_order=False)
    order.extend(list(reversed(expected)))
    lin = Linearization(objective, ordering=order)
    assert objective._engine_ordering == tuple(v.name for v in reversed(expected))
    assert [v.name for v in lin.ordering] == list(objective._engine_ordering)
    with pytest.raises(RuntimeError, match="CUDA"):      # no CPU engine: building it fails loudly, after the order has been accepted
        lin.engine
    Linearization(objective)
    assert objective._engine_ordering is None
    short = th.VariableOrdering(objective, default_order=False)
    short.append(expected[0])
    with pytest.raises(ValueError, match="not complete"):
        Linearization(objective, ordering=short)


def test_objective_and_cost_function_copies_share_nothing_but_names():
    """objective.py:643-700, theseus_function.py:90-108: copy() of an objective / cost function -- same names, dims and connectivity,
    independent tensors; a variable or weight used by several cost functions stays ONE object in the copy."""
    d = torch.float64
    a, b = th.SE3(name="a", dtype=d), th.SE3(name="b", dtype=d)
    z = th.SE3(name="z", dtype=d)
    w = th.DiagonalCostWeight(th.Variable(torch.ones(1, 6, dtype=d), name="w"), name="cw")
    obj = th.Objective(dtype=d)
    obj.add(th.Between(a, b, z, w, name="e1"))
    obj.add(th.Difference(a, th.SE3(name="t", dtype=d), w, name="p1"))
    obj.add(th.RobustCostFunction(th.Between(b, a, z, w, name="e2"), th.HuberLoss, th.Variable(torch.zeros(1, 1, dtype=d), name="r"), name="rob"))
    new = obj.copy()
    assert list(new.cost_functions) == list(obj.cost_functions) and list(new.optim_vars) == list(obj.optim_vars)
    assert sorted(new.aux_vars) == sorted(obj.aux_vars) and new.dim() == obj.dim()
    assert new.optim_vars["a"] is not a and new.optim_vars["a"].tensor.data_ptr() != a.tensor.data_ptr()
    assert new.cost_functions["e1"].optim_var_at(0) is new.cost_functions["p1"].optim_var_at(0) is new.optim_vars["a"]
    assert new.cost_functions["e1"].weight is new.cost_functions["p1"].weight and new.cost_functions["e1"].weight is not w
    assert new.cost_functions["rob"].cost_function.optim_var_at(1) is new.optim_vars["a"]
    assert new.cost_functions["e1"].aux_var_at(0) is new.aux_vars["z"] and new.cost_functions["e1"].num_aux_vars() == 1
    cf = obj.cost_functions["e1"].copy()
    assert cf.name == "e1_copy" and cf.optim_var_at(0).name == "a_copy" and cf.weight is not w
    cf.set_optim_var_at(0, b)
    assert cf.optim_var_at(0) is b and obj.cost_functions["e1"].optim_var_at(0) is a


def test_copy_of_an_autodiff_objective_is_independent():
    d = torch.float64
    x = th.Vector(1, name="x", dtype=d)
    a = th.Variable(torch.ones(3, 5, dtype=d), name="a")
    cf = th.AutoDiffCostFunction([x], lambda optim_vars, aux_vars: optim_vars[0].tensor * aux_vars[0].tensor, 5, aux_vars=[a], name="ad",
                                 cost_weight=th.ScaleCostWeight(torch.ones(1, dtype=d)))
    obj = th.Objective(dtype=d)
    obj.add(cf)
    new = obj.copy()
    ncf = new.cost_functions["ad"]
    assert new.aux_vars["a"] is ncf.aux_vars[0] and ncf.aux_vars[0] is not a
    ncf.aux_vars[0].tensor = torch.full((3, 5), 3.0, dtype=d)
    ncf.optim_vars[0].tensor = torch.full((3, 1), 2.0, dtype=d)
    assert float(ncf.error()[0, 0]) == 6.0 and float(cf.error()[0, 0]) == 0.0
