"""Host logic of the inference backbone copy (model.FoldedBackbone): folding BatchNorm into the
convolutions, the bias hand-over to the SE tail and the padded stem input are arithmetic identities -- checked here on the CPU
(where the copy runs on torch ops only) against the module it copies; the CUDA kernels behind it are checked in
tests/test_model_gpu.py."""
import torch

import rcnn_ocr_b200 as R


def _randomise_bn(cnn, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in cnn.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def test_folded_backbone_equals_the_module_on_the_host():
    torch.manual_seed(0)
    cnn = R.SEResNet31(3, 512).eval()
    _randomise_bn(cnn)
    folded = R.FoldedBackbone(cnn, torch.float32)
    x = torch.rand(2, 3, 32, 64) * 2 - 1
    with torch.no_grad():
        want, got = cnn(x), folded(x)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    # no parameters of its own in a state dict, nothing registered twice
    assert all(not k.startswith(("w", "b")) for k in folded.state_dict().keys() if k[1:].isdigit())
    assert folded.in_pad == 8 and getattr(folded, "w0").shape[1] == 8 and float(getattr(folded, "w0")[:, 3:].abs().max()) == 0.0


def test_rcnn_uses_the_folded_copy_in_eval_mode_only():
    torch.manual_seed(1)
    model = R.RCNN(20, hidden_size=64)
    _randomise_bn(model.cnn, 1)
    keys = set(model.state_dict().keys())
    model.eval().fold_backbone(torch.float32)
    assert set(model.state_dict().keys()) == keys            # the copy lives outside the module tree
    x = torch.rand(1, 3, 32, 64) * 2 - 1
    with torch.no_grad():
        a = model._features(x)
        model.fold_backbone(None)
        b = model._features(x)
    assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item()
    model.fold_backbone(torch.float32)
    model.train()
    assert getattr(model, "_folded", None) is not None       # kept, but train() mode goes through the module (batch statistics)
    with torch.no_grad():
        c = model._features(x)
    assert not torch.allclose(c, a)                          # BatchNorm batch statistics differ from the folded running ones
