"""Boundary behaviours of the reference's classes that need no GPU: mixtures of user-defined densities (the
per-point contract of pypmc/density/base.py:42-50), the assertion messages of ``MixtureDensity.multi_evaluate``
(pypmc/density/mixture_test.py:116-126), state after a failed ``update`` (pypmc/density/gauss_test.py:125-138,
student_t_test.py) and the component-death branch of the PMC update (pypmc/mix_adapt/pmc.pyx:227-244)."""
import logging

import numpy as np
import pytest

from pypmc_b200.density.base import ProbabilityDensity
from pypmc_b200.density.gauss import Gauss
from pypmc_b200.density.mixture import MixtureDensity, create_gaussian_mixture
from pypmc_b200.density.student_t import StudentT


class DummyComponent(ProbabilityDensity):
    """pypmc/density/mixture_test.py:15-23: evaluates to a constant, implements no multi_evaluate of its own."""

    def __init__(self, propose=[0.], eval_to=42.):
        self.to_propose = np.array(propose)
        self.dim = len(self.to_propose)
        self.eval_to = eval_to

    def evaluate(self, x):
        return self.eval_to

    def propose(self, N=1):
        return np.array([self.to_propose for i in range(N)])


TARGET = 39.69741490700607                     # mixture_test.py:33
MIX = MixtureDensity((DummyComponent(eval_to=10.), DummyComponent()), (.9, .1))


def test_user_density_mixture_evaluates_like_the_reference():
    at = np.array((-5.,))
    assert MIX.evaluate(at) == pytest.approx(TARGET, abs=1e-7)                       # mixture_test.py:79-80
    samples = np.array([at] * 2)
    individual = np.zeros((2, 2))
    out1, out2 = np.zeros(2), np.zeros(2)
    res1 = MIX.multi_evaluate(samples, individual=individual)
    res2 = MIX.multi_evaluate(samples, individual=individual, out=out1)
    res3 = MIX.multi_evaluate(samples, out=out2)
    assert res2 is out1 and res3 is out2
    for other in (res2, res3, out1, out2):                                           # bitwise, mixture_test.py:92-96
        np.testing.assert_equal(res1, other)
    np.testing.assert_array_almost_equal(res1, [TARGET] * 2)
    np.testing.assert_array_almost_equal(individual[:, 0], 10.)
    np.testing.assert_array_almost_equal(individual[:, 1], 42.)
    # components= fills only those columns and returns None (mixture.pyx:153-156)
    ind = np.full((2, 2), -1.0)
    assert MIX.multi_evaluate(samples, individual=ind, components=[1]) is None
    np.testing.assert_array_equal(ind, [[-1.0, 42.0]] * 2)


def test_error_messages_multi_evaluate():
    # mixture_test.py:105-126, same regular expressions
    samples = np.array([[1.], [2.], [3.]])
    samples_wrong_dim = np.array([[1., 1.2], [2., 32.], [2, 3.]])
    individual_ok = np.empty((3, 2))
    MIX.multi_evaluate(samples, individual=individual_ok)
    with pytest.raises(AssertionError, match='x.*wrong dim.*'):
        MIX.multi_evaluate(samples_wrong_dim, individual=individual_ok)
    with pytest.raises(AssertionError, match='individual.*must.*shape'):
        MIX.multi_evaluate(samples, individual=np.empty((2, 2)))
    with pytest.raises(AssertionError, match='individual.*must.*shape'):
        MIX.multi_evaluate(samples, individual=np.empty((3, 3)))
    with pytest.raises(AssertionError, match='components.*not None.*out.*must be None'):
        MIX.multi_evaluate(samples, np.empty(3), components=[0])
    with pytest.raises(AssertionError, match='out.*must.*len.*3'):
        MIX.multi_evaluate(samples, np.empty(9))
    with pytest.raises(TypeError):
        MIX.multi_evaluate(None)
    # the all-Gauss (CUDA) route raises the same messages before any launch
    gm = create_gaussian_mixture([[0.0], [1.0]], [[[1.0]], [[2.0]]])
    with pytest.raises(AssertionError, match='x.*wrong dim.*'):
        gm.multi_evaluate(samples_wrong_dim)
    with pytest.raises(AssertionError, match='individual.*must.*shape'):
        gm.multi_evaluate(samples, individual=np.empty((2, 2)))
    with pytest.raises(AssertionError, match='components.*not None.*out.*must be None'):
        gm.multi_evaluate(samples, np.empty(3), components=[0])
    with pytest.raises(AssertionError, match='out.*must.*len.*3'):
        gm.multi_evaluate(samples, np.empty(9))


def test_mixture_construction_contract():
    # mixture_test.py:38-77
    comps = [DummyComponent() for _ in range(5)]
    MixtureDensity(comps)
    comps[2].dim = 100
    with pytest.raises(AssertionError):
        MixtureDensity(comps)
    mix = MixtureDensity([DummyComponent for _ in range(5)])
    assert mix.normalized()
    np.testing.assert_allclose(mix.weights, 0.2, rtol=1e-15)
    mix.weights[0] = 2
    assert not mix.normalized()
    mix.normalize()
    assert mix.normalized()
    mix = MixtureDensity([DummyComponent for _ in range(5)], range(5))
    assert mix.prune() == [(0, DummyComponent, 0.)]
    assert len(mix.weights) == 4 and mix.normalized()


OFFDIAG = np.array([[0.01, 0.003], [0.003, 0.0025]])
SINGULAR = np.array([[0.0, 0.0], [0.0, 0.0025]])
ASYMMETRIC = np.array([[0.01, 0.002], [0.001, 0.0025]])


class _FakeRng(object):
    def normal(self, a, b, N):
        return np.array([0.7, -0.3][:N])

    def chisquare(self, dof):
        return 3.0


def test_state_unchanged_after_failed_update():
    # gauss_test.py:125-138 and its Student-t twin: LinAlgError must leave the component exactly as it was
    mean, point = np.array([4.3, 1.1]), np.array([4.35, 1.2])
    for comp, args in ((Gauss(mean, OFFDIAG), ()), (StudentT(mean, OFFDIAG, 5.0), (5.0,))):
        sample = comp.propose(1, _FakeRng())[0]
        record = comp._packed_record().copy()
        for bad in (SINGULAR, ASYMMETRIC):
            with pytest.raises(np.linalg.LinAlgError):
                comp.update(point, bad, *args)
        np.testing.assert_equal(comp.sigma, OFFDIAG)
        np.testing.assert_equal(comp.mu, mean)
        assert comp.dim == 2
        np.testing.assert_equal(comp.propose(1, _FakeRng())[0], sample)
        np.testing.assert_equal(comp._packed_record(), record)        # what the CUDA kernel would be given
    with pytest.raises(AssertionError, match=r'Dimensions of mean \(2\) and covariance matrix \(3\) do not match!'):
        Gauss(np.ones(2), np.eye(3))


def test_component_death_branch_of_the_update(caplog):
    """pmc.pyx:227-244: a component whose new covariance is not positive definite keeps its old parameters and gets
    weight zero; the caller then renormalises."""
    from pypmc_b200.mix_adapt.pmc import _apply_update
    mix = create_gaussian_mixture([[0.0, 0.0], [5.0, 5.0], [9.0, 1.0]], [np.eye(2), 2 * np.eye(2), OFFDIAG], [0.2, 0.3, 0.5])
    old = [(c.mu.copy(), c.sigma.copy()) for c in mix.components]
    alpha = np.array([0.5, 0.25, 0.25])
    mean = np.array([[1.0, 1.0], [4.0, 4.0], [8.0, 2.0]])
    cov = np.array([np.eye(2) * 3, SINGULAR, ASYMMETRIC * 0 + np.array([[1.0, 2.0], [2.0, 1.0]])])   # 1: singular, 2: indefinite
    with caplog.at_level(logging.WARNING, logger="pypmc_b200.mix_adapt.pmc"):
        failed = _apply_update(mix, [0, 1, 2], alpha, mean, cov)
    assert failed
    assert sum("Could not update component" in r.message for r in caplog.records) == 2
    np.testing.assert_equal(mix.components[0].mu, mean[0])
    np.testing.assert_equal(mix.components[0].sigma, cov[0])
    for k in (1, 2):
        np.testing.assert_equal(mix.components[k].mu, old[k][0])
        np.testing.assert_equal(mix.components[k].sigma, old[k][1])
        assert mix.weights[k] == 0.0
    assert mix.weights[0] == 0.5
    mix.normalize()
    assert mix.weights[0] == 1.0
    # Student-t: the degree of freedom is restored too (pmc.pyx:713-737)
    from pypmc_b200.density.mixture import create_t_mixture
    tm = create_t_mixture([[0.0, 0.0], [5.0, 5.0]], [np.eye(2), OFFDIAG], [3.0, 7.0], [0.5, 0.5])
    assert _apply_update(tm, [0, 1], np.array([0.6, 0.4]), mean[:2], np.array([np.eye(2), SINGULAR]), [4.0, 9.0])
    assert tm.components[0].dof == 4.0 and tm.components[1].dof == 7.0 and tm.weights[1] == 0.0
    np.testing.assert_equal(tm.components[1].sigma, OFFDIAG)


def test_device_samples_argument_contract():
    """Explicit weights / latent next to a DeviceSamples object are refused unless they are the objects it was built from
    (ADVICE r1); the rb / mincount checks look at the latent indices that will actually be used."""
    from pypmc_b200.mix_adapt import pmc
    ds = pmc.DeviceSamples.__new__(pmc.DeviceSamples)        # no GPU here: fill the fields the checks read
    lat = np.array([0, 1, 1])
    ds._src_weights, ds._src_latent, ds.latent = None, lat, lat
    assert pmc._as_device_samples(ds, None, None) is ds
    assert pmc._as_device_samples(ds, None, lat) is ds
    with pytest.raises(ValueError, match="weights"):
        pmc._as_device_samples(ds, np.ones(3), None)
    with pytest.raises(ValueError, match="latent"):
        pmc._as_device_samples(ds, None, np.array([0, 1, 1]))
    pmc._check_arguments(ds, None, ds.latent, 2, False)       # latent inside the DeviceSamples: rb=False / mincount are legal
    with pytest.raises(ValueError, match="rb"):
        pmc._check_arguments(np.zeros((3, 2)), None, None, 0, False)


def test_logsumexp2D_host_mirror():
    # tools/regularize_test.py:10-24
    from pypmc_b200.tools._regularize import logsumexp2D
    vals = np.array([[1., 2., 3.], [4., 5., 6.]])
    ref = np.log((np.array([1.3, 0.4, 0.3]) * np.exp(vals)).sum(1))
    np.testing.assert_allclose(logsumexp2D(vals, np.array([1.3, 0.4, 0.3])), ref, rtol=1e-15)
    with pytest.raises(AssertionError, match="negative weight"):
        logsumexp2D(vals, np.array([1.0, -0.1, 0.1]))
