"""engine.Grads (CPU, no kernels): every .grad is a view of ONE persistent flat buffer -- the buffer a data-parallel run
all-reduces with a single collective and whose addresses the captured backward graph of graph mode relies on."""
import torch


def make():
    from nero_b200.engine import Grads
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2, 2))]
    return Grads(ps), ps


def test_views_are_persistent_across_zero_grad():
    g, ps = make()
    flat = g.ensure()
    assert flat.numel() == 12 + 5 + 4 and all(p.grad is not None for p in ps)
    ptrs = [p.grad.data_ptr() for p in ps]
    ps[0].grad.add_(1.0)
    assert float(flat[:12].sum()) == 12.0               # views alias the flat buffer, in parameter order
    for p in ps:
        p.grad = None                                   # zero_grad(set_to_none=True)
    flat2 = g.ensure()
    assert flat2 is flat and [p.grad.data_ptr() for p in ps] == ptrs, 'same buffer, same addresses'
    assert float(flat.abs().sum()) == 0.0, 'a re-attached gradient starts from zero'


def test_partial_reattach_keeps_accumulated_gradients():
    g, ps = make()
    g.ensure()
    ps[0].grad.fill_(2.0)
    ps[1].grad.fill_(3.0)
    ps[1].grad = None
    g.ensure()
    assert float(ps[0].grad.sum()) == 24.0 and float(ps[1].grad.abs().sum()) == 0.0


def test_strict_moves_foreign_gradients_into_the_flat_buffer():
    g, ps = make()
    g.ensure()
    ps[2].grad = torch.full((2, 2), 7.0)                # somebody else's backward allocated this one
    flat = g.ensure(strict=True)
    assert ps[2].grad.data_ptr() == g.views[2].data_ptr()
    assert float(flat[-4:].sum()) == 28.0
