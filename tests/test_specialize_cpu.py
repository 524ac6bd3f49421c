"""Run-time specialisation of the element-wise kernels (waveome_b200/specialize.py): the generated CUDA text of several
kernel structures compiles with NVRTC for sm_100a (no GPU needed), the text depends on the structure only, and programs
the generator does not cover are declined."""
import numpy as np

import helpers
import waveome_b200 as wb
from waveome_b200 import engine, specialize


def _mixed_kernel():
    cat = wb.Categorical(active_dims=[3]); wb.set_trainable(cat.variance, False)
    se_frozen = wb.SquaredExponential(active_dims=[2], lengthscales=0.7); wb.set_trainable(se_frozen.lengthscales, False)
    return wb.Sum([
        wb.Categorical(active_dims=[0]),
        wb.Constant(variance=0.3),
        wb.Lin(active_dims=[2], variance=0.5),
        wb.Matern12(active_dims=[1]), wb.Matern32(active_dims=[2], lengthscales=0.7), wb.Matern52(active_dims=[1], lengthscales=1.3),
        wb.Periodic(wb.SquaredExponential(active_dims=[1]), period=1.7),
        wb.Product([cat, wb.SquaredExponential(active_dims=[1], lengthscales=0.8)]),
        wb.Product([wb.SquaredExponential(active_dims=[1]), wb.SquaredExponential(active_dims=[2])]),
        wb.Product([wb.Lin(active_dims=[1]), wb.Matern32(active_dims=[2])]),
        wb.Product([cat, wb.Periodic(wb.SquaredExponential(active_dims=[2]), period=0.9), se_frozen]),
    ])


def test_generated_text_compiles_for_sm100a():
    for k in (helpers.saturated_kernel(), _mixed_kernel()):
        model = wb.GPR(k, mean_function=wb.ConstantMean(0.0))
        src = specialize.generate(model.program())
        assert src is not None and src.gram_name in src.source and src.grad_name in src.source
        assert engine.rtc_check(src.source) > 10000          # cubin bytes
        assert src.gram_smem < 227 * 1024 and src.grad_smem < 227 * 1024


def test_text_depends_on_structure_only():
    a = wb.GPR(helpers.saturated_kernel(hs=1.0), mean_function=wb.ConstantMean(0.0))
    b = wb.GPR(helpers.saturated_kernel(hs=7.0), mean_function=wb.ConstantMean(0.5), noise_variance=0.2)
    sa, sb = specialize.generate(a.program()), specialize.generate(b.program())
    assert sa.key == sb.key                                  # priors and values are read from the device program
    c = wb.GPR(helpers.saturated_kernel(num=(2, 1)), mean_function=wb.ConstantMean(0.0))
    assert specialize.generate(c.program()).key != sa.key
    frozen = helpers.saturated_kernel()
    wb.set_trainable(frozen.kernels[2].lengthscales, False)
    assert specialize.generate(wb.GPR(frozen, mean_function=wb.ConstantMean(0.0)).program()).key != sa.key


def test_uncovered_leaves_are_declined():
    m = wb.GPR(helpers.all_leaf_kernel(), mean_function=wb.ConstantMean(0.0))      # contains a polynomial leaf
    assert specialize.generate(m.program()) is None


def test_compile_error_is_reported():
    import pytest
    with pytest.raises(engine.EngineError):
        engine.rtc_check("__global__ void k() { this is not CUDA }")
