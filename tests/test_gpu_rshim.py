"""The reference-side binding itself: r_shim/signer_gibbs_shim.cpp -- the file that replaces src/gibbs_2.cpp and the
GibbsSamplerCpp half of src/RcppExports.cpp -- compiled UNMODIFIED against a stand-in Rcpp (oracle/r_stub, R is absent
from the image) and executed through its 24-SEXP entry `_signeR_GibbsSamplerCpp` (src/signeR_init.c:13).  Its 8-slot list
must equal, bit for bit, the list the reference's own src/gibbs_2.cpp returns (oracle/_ref) and the oracle's samples."""
import numpy as np
import pytest

from helpers import assert_bits_equal
from signer_b200 import synth

rshim = pytest.importorskip("oracle.rshim_runner")


def _inputs(I=24, J=45, K=4, seed=3):
    M, W, _, _ = synth.cohort(I, J, K, burden=30.0 * I, seed=seed)
    M[2, 5] = 3000.0
    st = synth.init_state(I, J, K, seed=seed + 1)
    return M, W, st


def test_shim_rejects_bad_shapes_without_a_gpu():
    if not rshim.available():
        pytest.skip("oracle/_ref/librshim.so not built")
    M, W, st = _inputs()
    I, J = M.shape
    K = st["P"].shape[1]
    Z = np.zeros((I, J, K), order="F")
    with pytest.raises(rshim.RError, match="dimensions"):      # Rcpp::stop -> END_RCPP -> R error, no device touched
        rshim.call(M, W[:, :-1], Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), 1, 1)
    assert rshim.seed_for(42) == rshim.seed_for(42) != rshim.seed_for(43)


@pytest.mark.gpu
@pytest.mark.parametrize("keep_par,Pfixed", [(True, False), (False, False), (True, True)])
def test_shim_list_equals_reference_source_and_oracle(oracle, keep_par, Pfixed):
    if not rshim.available():
        pytest.skip("oracle/_ref/librshim.so not built")
    from oracle import ref_runner
    M, W, st = _inputs()
    I, J = M.shape
    K = st["P"].shape[1]
    h = synth.hyper_tuple()
    hd = dict(zip(("ap", "bp", "ae", "be", "lp", "le", "var_ap", "var_ae"), h))
    r_seed = 20241020
    seed = rshim.seed_for(r_seed)
    # R builds the initial Z itself (R/cpp_eBayesNMF.R:79-89) and passes the cube: use the oracle's initial allocation
    Z0 = oracle.gibbs_run(M, W, None, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=hd, burn=0, eval=0, seed=seed)["Z"]
    got = rshim.call(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], h, 6, 4, Pfixed=Pfixed, keep_par=keep_par,
                     r_seed=r_seed)
    want = oracle.gibbs_run(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=hd, burn=6, eval=4,
                            Pfixed=Pfixed, keep_par=keep_par, seed=seed)
    names = {0: "Z", 2: "Ps", 3: "Es", 4: "Aps", 5: "Bps", 6: "Aes", 7: "Bes"}
    for slot, name in names.items():
        if got[slot] is None:
            assert not keep_par and slot in (0, 4, 6)
            continue
        g = got[slot][:, :, :, 0] if slot == 0 else got[slot]
        assert_bits_equal(g, want[name], "shim slot %d vs oracle %s" % (slot + 1, name))
    if ref_runner.available():                                   # ... and the reference's own source, same inputs
        ref = ref_runner.gibbs_run(M, W, Z0, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], hyper=hd, burn=6, eval=4,
                                   Pfixed=Pfixed, keep_par=keep_par, seed=seed)
        assert ref["list_shape_ok"]
        for slot, name in names.items():
            if got[slot] is None:
                continue
            g = got[slot][:, :, :, 0] if slot == 0 else got[slot]
            assert_bits_equal(g, ref[name], "shim slot %d vs reference source %s" % (slot + 1, name))


@pytest.mark.gpu
def test_shim_turns_library_errors_into_r_errors():
    if not rshim.available():
        pytest.skip("oracle/_ref/librshim.so not built")
    M, W, st = _inputs()
    I, J = M.shape
    K = st["P"].shape[1]
    Z = np.zeros((I, J, K), order="F")
    with pytest.raises(rshim.RError, match="not reachable|not implemented"):     # Zfixed = TRUE (src/gibbs_2.cpp:376-405)
        rshim.call(M, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), 1, 1, Zfixed=True)
    Mbad = M.copy(); Mbad[0, 0] = -1.0
    with pytest.raises(rshim.RError, match="negative"):
        rshim.call(Mbad, W, Z, st["P"], st["E"], st["Ap"], st["Bp"], st["Ae"], st["Be"], synth.hyper_tuple(), 1, 1)
