"""Final 1x1 classifier conv fused into the distance head (``dml_conv1x1_head_forward``)
against torch's conv + the CPU oracle's distance block (anomaly/models/models.py:609,636-657): embedding <= 1e-5
relative to its scale, logits <= 1e-5 relative (the north-star tolerance for distances)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,c,k,h,w,bias", [(2, 512, 13, 38, 67, True), (1, 256, 16, 64, 128, True), (3, 100, 17, 9, 11, False),
                                            (1, 512, 19, 71, 125, True), (2, 7, 1, 5, 3, True), (1, 130, 32, 16, 16, False)])
def test_conv_head_matches_conv_plus_oracle(b, c, k, h, w, bias):
    from dml_b200 import head as H
    g = torch.Generator().manual_seed(c + k)
    f = torch.relu(torch.randn(b, c, h, w, generator=g))
    conv = nn.Conv2d(c, k, 1, bias=bias)
    with torch.no_grad():
        conv.weight.mul_(0.5)
        emb_ref = conv.double()(f.double()).float()      # float64 ground truth of the contraction
    z_ref = O.distance_logits(emb_ref, O.make_centers(k))
    emb, z = H.conv1x1_head(f.cuda(), conv.weight.float(), conv.bias.float() if bias else None, 3.0)
    scale = float(emb_ref.abs().max())
    assert float((emb.cpu() - emb_ref).abs().max()) <= 1e-5 * scale
    zs = z_ref.abs().amax(1, keepdim=True)
    assert float(((z.cpu() - z_ref).abs() / zs).max()) <= 1e-5
    # outputs can be requested separately
    e2, z2 = H.conv1x1_head(f.cuda(), conv.weight.float(), conv.bias.float() if bias else None, 3.0, want_embedding=False)
    assert e2 is None and torch.equal(z2, z)


def test_ppm_decoder_lowres_uses_the_fused_conv_head():
    """PPMDeepsup_embedding.forward_lowres in eval mode: fused conv + head == the unfused module path"""
    from dml_b200.anomaly.models import PPMDeepsup_embedding
    torch.backends.cudnn.allow_tf32 = False          # the comparison path must be fp32 too (cuDNN convs default to TF32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    dec = PPMDeepsup_embedding(num_class=13, fc_dim=64, use_softmax=True).cuda().eval()
    conv_out = [torch.randn(1, 32, 12, 20).cuda(), torch.randn(1, 64, 12, 20).cuda()]
    with torch.no_grad():
        z, emb = dec.forward_lowres(conv_out)
        conv5 = conv_out[-1]
        ppm = torch.cat([conv5] + [nn.functional.interpolate(p(conv5), conv5.shape[2:], mode='bilinear', align_corners=False)
                                   for p in dec.ppm], 1)
        emb_ref = dec.conv_last(ppm)
    np.testing.assert_allclose(emb.cpu().numpy(), emb_ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    z_ref = O.distance_logits(emb_ref.cpu(), O.make_centers(13))
    np.testing.assert_allclose(z.cpu().numpy(), z_ref.numpy(), rtol=1e-4, atol=1e-4)


def test_evaluate_image_fused_equals_the_reference_loop():
    """anomaly.eval_ood.evaluate_image (stride-8 logits per scale + ONE fused upsample / average / score kernel) against
    the reference's loop replayed with the same modules: forward(segSize) per scale, scores += ./n, score lines."""
    from types import SimpleNamespace
    from dml_b200.anomaly.eval_ood import evaluate_image, score_map
    from dml_b200.anomaly.models import PPMDeepsup_embedding
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.c4 = nn.Conv2d(3, 32, 3, stride=8, padding=1)
            self.c5 = nn.Conv2d(3, 64, 3, stride=8, padding=1)

        def forward(self, x, return_feature_maps=False):
            return [self.c4(x), self.c5(x)]

    dec = PPMDeepsup_embedding(num_class=13, fc_dim=64, use_softmax=True)
    with torch.no_grad():
        dec.conv_last[4].weight.mul_(3.0)
    module = SimpleNamespace(encoder=Enc().cuda().eval(), decoder=dec.cuda().eval(),
                             parameters=lambda: iter(dec.parameters()))
    H, W = 96, 160
    imgs = [torch.randn(1, 3, h, w) for (h, w) in [(40, 72), (56, 96), (72, 120)]]
    batch = {"img_data": imgs, "seg_label": torch.zeros(1, H, W, dtype=torch.long)}
    cfg = SimpleNamespace(OOD=SimpleNamespace(ood="dissum", exclude_back=False))
    pred, conf = evaluate_image(module, batch, cfg)
    with torch.no_grad():
        scores = torch.zeros(1, 13, H, W, device="cuda")
        for img in imgs:
            s_tmp, _ = module.decoder(module.encoder(img.cuda(), return_feature_maps=True), segSize=(H, W))
            scores = scores + s_tmp / len(imgs)
        pred_ref, conf_ref = score_map(scores, "dissum", False)
    assert (pred.cpu() != pred_ref[0].cpu()).float().mean() < 1e-3
    np.testing.assert_allclose(conf.cpu().numpy(), conf_ref[0].cpu().numpy(), rtol=1e-4, atol=1e-4)
