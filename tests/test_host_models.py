"""Host-side mirror of the reference's model constructors (no GPU, no kernels): constructor / factory parity of the
DeepLab DML wrappers with DeepLabV3Plus-Pytorch/network/utils.py:56-60,121-135 and network/modeling.py:140-158."""
import pytest
import torch.nn as nn


class _Head(nn.Module):
    def __init__(self, inplanes, low_level_planes, num_classes, aspp_dilate):
        super().__init__()
        self.args = (inplanes, low_level_planes, num_classes, tuple(aspp_dilate))


def test_self_distillation_constructor_parity():
    from dml_b200.deeplab.network import _SimpleSegmentationModel_embedding_self_distillation as M
    m = M(nn.Identity(), head_factory=_Head)                       # the reference's call shape: cls(backbone)
    assert m.classifier_list == ["classifier", "classifier_1"] and m.cls_novel == 1
    assert m.classifier.args == (2048, 256, 16, (6, 12, 18)) and m.classifier_1.args == (2048, 256, 17, (6, 12, 18))
    assert tuple(m.centers.shape) == (17, 17)
    # the optimiser of test_self_distillation.py:476-478 addresses model.classifier_1
    assert hasattr(m, "classifier_1")
    m2 = M(nn.Identity(), [nn.Identity(), nn.Identity(), nn.Identity()])
    assert m2.classifier_list == ["classifier", "classifier_1", "classifier_2"]


def test_factories_build_on_the_reference_tree():
    """deeplabv3plus_embedding[_self_distillation]_resnet101 inside the reference's package layout (the copy under
    baseline/_ref or /root/reference): reference backbone + heads, this repo's wrappers."""
    from oracle import ref_loader
    if ref_loader.reference_root() is None:
        pytest.skip("no reference tree available (baseline/_ref not installed)")
    from dml_b200.deeplab import network as N
    with ref_loader.reference("DeepLabV3Plus-Pytorch"):
        m = N.deeplabv3plus_embedding_self_distillation_resnet101(num_classes=16, output_stride=16, pretrained_backbone=False)
        e = N.deeplabv3plus_embedding_resnet101(num_classes=16, output_stride=16, pretrained_backbone=False)
        import network as ref_net
        ref = ref_net.deeplabv3plus_embedding_self_distillation_resnet101(num_classes=16, output_stride=16, pretrained_backbone=False)
    assert isinstance(m, N._SimpleSegmentationModel_embedding_self_distillation)
    assert isinstance(e, N._SimpleSegmentationModel_embedding)
    # same parameter names / shapes as the reference model: checkpoints load unchanged
    a = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
