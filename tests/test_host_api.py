"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
the headers declare, fails loudly without a GPU, and the rnnoise.h shim exports the reference's
C++-mangled names.  No compute."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    from percepnet_b200 import build
    build.build()
    return build


def test_library_exports_every_declared_symbol(built):
    from percepnet_b200 import api
    hdr = open(os.path.join(ROOT, "include", "percepnet_b200.h")).read()
    declared = set(re.findall(r"\b(pnb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pnb_model", "pnb_engine"}
    assert declared == set(api.EXPORTS)
    L = api.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.pnb_version()


def test_layout_structs_match_header():
    """ctypes mirrors == include/pnb_nnet_layout.h == reference nnet.h:44-89 field order."""
    from percepnet_b200 import weights as W
    assert [f[0] for f in W.DenseLayerC._fields_] == ["bias", "input_weights", "nb_inputs", "nb_neurons", "activation"]
    assert [f[0] for f in W.GRULayerC._fields_] == ["bias", "input_weights", "recurrent_weights", "nb_inputs",
                                                    "nb_neurons", "activation", "reset_after"]
    assert [f[0] for f in W.Conv1DLayerC._fields_] == ["bias", "input_weights", "nb_inputs", "kernel_size",
                                                       "nb_neurons", "activation"]
    assert [f[0] for f in W.ModelC._fields_] == ["fc", "conv1", "conv2", "gru1", "gru2", "gru3", "gru_gb", "gru_rb",
                                                 "fc_gb", "fc_rb"]
    assert C.sizeof(W.ModelC) == 80 and C.sizeof(W.GRULayerC) == 40 and C.sizeof(W.DenseLayerC) == 32


def test_create_fails_loudly_without_gpu(built, model0):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from percepnet_b200 import api
    with pytest.raises(api.PnbError) as ei:
        api.Engine(4, 4, model0)
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)


def test_rejects_wrong_architecture(built, model0):
    from percepnet_b200 import api
    L = api.load_library()
    m = model0.as_c_model()
    bad = type(m)()
    C.memmove(C.byref(bad), C.byref(m), C.sizeof(m))
    fc = type(m.fc.contents)()
    C.memmove(C.byref(fc), m.fc, C.sizeof(fc))
    fc.nb_neurons = 64
    bad.fc = C.pointer(fc)
    h = C.c_void_p()
    assert L.pnb_create(C.byref(h), 1, 1, C.byref(bad), 0, 0) == -1
    assert b"PercepNet" in L.pnb_last_error()


def test_rnnoise_shim_exports_reference_symbols(built):
    """The names an unmodified src/main.cpp imports (PROBE nm)."""
    out = subprocess.run(["nm", "-D", "--defined-only", built.SHIM], capture_output=True, text=True, check=True).stdout
    for sym in ("_Z14rnnoise_createP8RNNModel", "_Z21rnnoise_process_frameP12DenoiseStatePfPKfP8_IO_FILE",
                "_Z15rnnoise_destroyP12DenoiseState", "_Z16rnnoise_get_sizev", "_Z12rnnoise_initP12DenoiseStateP8RNNModel"):
        assert sym in out, sym


def test_synthetic_weights_are_reproducible(model0):
    from percepnet_b200.weights import N_PARAMS, synth_model
    assert model0.n_params() == N_PARAMS
    assert synth_model(0).digest() == model0.digest()
    assert synth_model(1).digest() != model0.digest()
    w = model0.arrays["gru1_weights"]
    assert w.shape == (512, 1536) and abs(float(w.std()) - (1 / np.sqrt(512)) / np.sqrt(3)) < 1e-3


def test_weight_blob_roundtrip(built, model0, tmp_path):
    """f4: PackedModel.save_blob -> pnb_model_load_blob gives back the same arrays in the RNNModel layout; a
    truncated or foreign file is refused."""
    from percepnet_b200 import api
    from percepnet_b200.weights import ModelC
    path = str(tmp_path / "w.pnbw")
    model0.save_blob(path)
    assert os.path.getsize(path) < 33_000_000
    bm = api.BlobModel(path)
    m = C.cast(bm.ptr, C.POINTER(ModelC)).contents
    ref = model0.as_c_model()
    for name, n in (("fc", 70 * 128), ("conv2", 3 * 512 * 512), ("fc_gb", 2560 * 34)):
        a = np.ctypeslib.as_array(getattr(m, name).contents.input_weights, (n,))
        b = np.ctypeslib.as_array(getattr(ref, name).contents.input_weights, (n,))
        assert np.array_equal(a, b)
    g = m.gru_rb.contents
    assert (g.nb_inputs, g.nb_neurons, g.reset_after) == (1024, 128, 1)
    assert np.array_equal(np.ctypeslib.as_array(g.recurrent_weights, (128 * 384,)), model0.arrays["gru_rb_recurrent_weights"].ravel())
    assert np.array_equal(np.ctypeslib.as_array(g.bias, (768,)), model0.arrays["gru_rb_bias"])
    bm.free()
    open(path, "r+b").truncate(1_000_000)
    with pytest.raises(api.PnbError):
        api.BlobModel(path)
    out = subprocess.run(["nm", "-D", "--defined-only", built.SHIM], capture_output=True, text=True, check=True).stdout
    assert "_Z23rnnoise_model_from_fileP8_IO_FILE" in out and "_Z18rnnoise_model_freeP8RNNModel" in out


def test_public_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both headers must compile as strict C99 and as C++11 (no torch, no CUDA types)."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "percepnet_b200.h"\n#include "pnb_nnet_layout.h"\n'
                   "int main(void) { pnb_engine *e = 0; (void)e; return (int)sizeof(pnb_model) > 0 ? 0 : 1; }\n")
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    subprocess.run([cxx, "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True)
    hdr = open(os.path.join(inc, "percepnet_b200.h")).read()
    assert "#include <torch" not in hdr and "cuda_runtime" not in hdr and "at::" not in hdr
