"""Genome-column sharding (signer_gibbs_run_sharded): the sharded chain must equal the oracle run with the same
number of shards bit for bit -- with the shards emulated on one GPU and, when several GPUs are visible, with a
real NCCL all-reduce between them."""
import numpy as np
import pytest

from helpers import assert_bits_equal, compare_all
from signer_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _run(sg, oracle, devices, I=24, J=520, K=5, burn=8, ev=4, keep_par=True, Pfixed=False, Z=None, seed=9):
    M, W, _, _ = synth.cohort(I, J, K, burden=25.0 * I, seed=J)
    M[1, 3] = 4000.0
    st = synth.init_state(I, J, K, seed=5)
    h = synth.hyper_tuple()
    got = sg.GibbsSamplerCpp(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], *h, burn, ev,
                             Pfixed, False, False, False, keep_par, seed=seed, shard_devices=devices)
    hd = dict(zip(("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae"), h))
    want = oracle.gibbs_run(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=hd, burn=burn, eval=ev,
                            Pfixed=Pfixed, keep_par=keep_par, seed=seed, nshards=len(devices))
    return got, want


@pytest.mark.parametrize("nsh", [1, 2, 3, 5])
def test_sharded_on_one_device_equals_oracle(sg, oracle, nsh):
    got, want = _run(sg, oracle, [0] * nsh)
    compare_all(got, want)


def test_sharded_keep_par_false_and_given_cube(sg, oracle):
    I, J, K = 24, 520, 5
    M, _, _, _ = synth.cohort(I, J, K, burden=25.0 * I, seed=J)
    M[1, 3] = 4000.0
    Z = np.zeros((I, J, K), order="F"); Z[:, :, 2] = np.trunc(M)
    got, want = _run(sg, oracle, [0, 0], keep_par=False, Z=Z)
    compare_all(got, want, keep_par=False)


def test_sharding_changes_only_rounding(sg, oracle):
    """Different shard counts are different (fixed) summation orders of W E': the integer allocations of the FIRST
    sweep's step 1 are identical, later values differ at rounding level only."""
    a, _ = _run(sg, oracle, [0], burn=0, ev=1)
    b, _ = _run(sg, oracle, [0, 0, 0], burn=0, ev=1)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-9)
    np.testing.assert_allclose(a[3], b[3], rtol=1e-9)


def test_sharded_over_nccl_equals_oracle(sg, oracle):
    nd = _lib.lib().signer_device_count()
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    devs = list(range(min(nd, 4)))
    got, want = _run(sg, oracle, devs, J=128 * 9 + 17, burn=10, ev=5)
    compare_all(got, want)
    emu, _ = _run(sg, oracle, [0] * len(devs), J=128 * 9 + 17, burn=10, ev=5)
    for slot in (2, 3, 4, 5, 6, 7):
        assert_bits_equal(got[slot], emu[slot], "slot %d nccl vs emulated" % slot)
