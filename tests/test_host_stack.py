"""CPU: host-side logic of the Stack mirror -- recording bookkeeping and the reference's
pre-checks (thrown before any device work, exactly where the reference throws them)."""
import numpy as np
import pytest


def test_new_recording_pushes_the_sentinel(pkg):
    s = pkg.Stack()
    assert s.n_statements() == 1 and s.n_operations() == 0       # Stack.h:528-543
    assert tuple(s.arrays()["stmt"][0]) == (-1, 0)
    x = s.register_gradients(3)
    assert x == 0 and s.max_gradients() == 3
    s.new_recording()
    assert s.max_gradients() == 4                                 # max_gradient_ = i_gradient_ + 1


def test_push_api_builds_the_reference_layout(pkg):
    s = pkg.Stack()
    s.register_gradients(4)
    s.new_recording()
    s.push_rhs(2.0, 0); s.push_rhs(3.0, 1); s.push_lhs(2)
    s.push_lhs(3)                                                 # x = constant: zero operations
    a = s.arrays()
    assert a["stmt"].tolist() == [[-1, 0], [2, 2], [3, 2]]
    assert a["mult"].tolist() == [2.0, 3.0] and a["idx"].tolist() == [0, 1]


def test_prechecks_throw_like_the_reference(pkg):
    s = pkg.Stack()
    s.register_gradients(2); s.new_recording()
    s.push_rhs(1.0, 0); s.push_lhs(1)
    with pytest.raises(pkg.gradients_not_initialized):           # adept/Stack.cpp:161-163
        s.compute_adjoint()
    with pytest.raises(pkg.gradients_not_initialized):           # adept/Stack.cpp:187-189
        s.compute_tangent_linear()
    with pytest.raises(pkg.gradients_not_initialized):           # Stack.h:311-313
        s.get_gradients(0, 1)
    with pytest.raises(pkg.dependents_or_independents_not_identified):   # adept/jacobian.cpp:208-210
        s.jacobian()
    s.independent(0)
    with pytest.raises(pkg.dependents_or_independents_not_identified):
        s.jacobian_reverse()
    with pytest.raises(pkg.gradient_out_of_range):               # Stack.h:296-298
        s.set_gradients(0, 99, np.zeros(99))
    s.set_gradients(0, 2, [1.0, 2.0])
    assert s.get_gradients(0, 2).tolist() == [1.0, 2.0]
    s.clear_gradients()
    with pytest.raises(pkg.gradients_not_initialized):
        s.get_gradients(0, 1)
    s.dependent(1)
    with pytest.raises(pkg.size_mismatch):                       # adept/jacobian.cpp:652-655
        s.jacobian_into(np.zeros((3, 3)))


def test_independent_dependent_lists(pkg):
    s = pkg.Stack()
    s.independent([0, 1, 2]); s.independent(5)
    s.dependent(np.array([7, 8]))
    assert s.n_independent() == 4 and s.n_dependent() == 2
    s.clear_independents(); s.clear_dependents()
    assert s.n_independent() == 0 and s.n_dependent() == 0
    assert s.max_jacobian_threads() == 1 and s.set_max_jacobian_threads(8) == 1
