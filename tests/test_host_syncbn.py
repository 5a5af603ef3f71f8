"""CPU tier: the host side of anomaly.lib.nn -- constructor / buffer parity with the reference's
_SynchronizedBatchNorm (anomaly/lib/nn/modules/batchnorm.py:38-55), the PyTorch path in evaluation / non-parallel mode
(:58-62), model conversion, and no CPU fallback for the synchronised kernels."""
import pytest
import torch
import torch.nn as nn


def test_constructor_and_buffers_mirror_the_reference():
    from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d
    m = SynchronizedBatchNorm2d(6)
    assert m.eps == 1e-5 and m.momentum == 0.001 and m.affine                      # the reference's defaults
    names = dict(m.named_buffers())
    for k in ("running_mean", "running_var", "_tmp_running_mean", "_tmp_running_var", "_running_iter"):
        assert k in names
    assert torch.equal(m._tmp_running_mean, torch.zeros(6)) and torch.equal(m._tmp_running_var, torch.ones(6))
    assert float(m._running_iter) == 1.0 and m._moving_average_fraction == pytest.approx(0.999)
    assert set(dict(m.named_parameters())) == {"weight", "bias"}
    with pytest.raises(ValueError):
        SynchronizedBatchNorm1d(4, always_sync=True).train()._check_input_dim(torch.zeros(2, 4, 3, 3))
    with pytest.raises(ValueError):
        SynchronizedBatchNorm3d(4)._check_input_dim(torch.zeros(2, 4, 3, 3))


def test_eval_and_non_parallel_training_use_the_pytorch_path():
    from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d
    torch.manual_seed(0)
    x = torch.randn(3, 5, 4, 6)
    ours, ref = SynchronizedBatchNorm2d(5), nn.BatchNorm2d(5, momentum=0.001)
    for train in (True, False):                                                     # one rank, no always_sync: F.batch_norm
        ours.train(train)
        ref.train(train)
        torch.testing.assert_close(ours(x), ref(x))
    torch.testing.assert_close(ours.running_mean, ref.running_mean)
    torch.testing.assert_close(ours.running_var, ref.running_var)


def test_no_cpu_fallback_for_the_synchronised_kernels():
    from dml_b200 import DmlError
    from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d
    m = SynchronizedBatchNorm2d(4, always_sync=True).train()
    with pytest.raises(DmlError):
        m(torch.randn(2, 4, 3, 3))


def test_convert_model_and_replication_callback():
    from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, convert_model, patch_replication_callback
    net = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4), nn.ReLU(), nn.Sequential(nn.Conv2d(4, 2, 1), nn.BatchNorm2d(2)),
                        nn.Flatten(), nn.BatchNorm1d(2 * 5 * 5))
    with torch.no_grad():
        net[1].running_mean.fill_(0.25)
        net[1].weight.fill_(2.0)
    conv_w = net[0].weight
    out = convert_model(net)
    assert out is net and isinstance(net[1], SynchronizedBatchNorm2d) and isinstance(net[3][1], SynchronizedBatchNorm2d)
    assert isinstance(net[5], SynchronizedBatchNorm1d) and net[0].weight is conv_w
    assert torch.equal(net[1].running_mean, torch.full((4,), 0.25)) and torch.equal(net[1]._tmp_running_mean, torch.full((4,), 0.25))
    assert torch.equal(net[1].weight, torch.full((4,), 2.0))
    net.eval()
    assert net(torch.randn(2, 3, 5, 5)).shape == (2, 50)
    assert patch_replication_callback(net) is net
