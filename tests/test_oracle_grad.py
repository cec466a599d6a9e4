"""Pin the GRADIENT oracle (torch CPU autograd through oracle/flow_oracle.py in train mode) against the gradient goldens
produced by the reference's own training-mode backward (tests/golden/make_golden_grad.py).  CPU-only."""
import pytest
import torch

from tests import _golden, _gradcheck as GC

torch.set_num_threads(4)

LAYER_CASES = [n for n in _golden.names('grad_') if not n.startswith('grad_model_')]
MODEL_CASES = _golden.names('grad_model_')


@pytest.mark.parametrize('name', LAYER_CASES)
def test_layer_gradients_match_reference(name):
    meta, a, _ = _golden.load(name)
    _, src, sd = _golden.load(meta['source'])
    z, l, gx, gl, grads = GC.oracle_layer_grads(meta, sd, src['x'], src['ldj0'], a['Rz'], a['Rl'])
    GC.grad_close(z, a['fwd_z'], 1e-5, 'z')
    GC.grad_close(l, a['fwd_ldj'], 1e-5, 'ldj')
    GC.grad_close(gx, a['gx'], 2e-5, 'gx')
    GC.grad_close(gl, a['gldj'], 1e-6, 'gldj')
    want = {k[5:]: v for k, v in a.items() if k.startswith('grad/')}
    assert set(want) <= set(grads), sorted(set(want) - set(grads))  # (buffers such as log_gamma with affine=False carry no gradient in the reference)
    floor = GC.grad_floor(want)
    for k, v in want.items():
        GC.grad_close(grads[k], v, 5e-5, k, floor)


@pytest.mark.parametrize('name', MODEL_CASES)
def test_model_gradients_match_reference(name):
    meta, a, _ = _golden.load(name)
    _, src, sd = _golden.load(meta['source'])
    loss, grads = GC.oracle_model_grads(meta, sd, src['x'])
    assert abs(float(loss) - float(a['loss'])) <= 2e-6 * abs(float(a['loss']))
    want = {k[5:]: v for k, v in a.items() if k.startswith('grad/')}
    assert set(want) <= set(grads), sorted(set(want) - set(grads))  # (buffers such as log_gamma with affine=False carry no gradient in the reference)
    floor = GC.grad_floor(want)
    for k, v in want.items():
        GC.grad_close(grads[k], v, 2e-4, k, floor)
