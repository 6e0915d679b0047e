import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_train as T
from unipose_b200 import train
for variant in ("random_masks", "ones_masks"):
    m, sd, x, target, masks = T._setup()
    if variant == "ones_masks":
        masks = [torch.ones_like(t) for t in masks]
    heat = train.forward_train(m, x.cuda(), dropout_masks=[t.cuda() for t in masks])
    loss = F.mse_loss(heat, target.cuda()); loss.backward()
    ref_heat, ref_loss, ref_sd = T._oracle_step(sd, x, target, masks)
    params = dict(m.named_parameters())
    keys = ["decoder.last_conv.8.weight", "decoder.last_conv.5.bias", "decoder.last_conv.5.weight", "decoder.last_conv.4.weight",
            "decoder.last_conv.1.bias", "decoder.last_conv.1.weight", "decoder.last_conv.0.weight", "decoder.bn1.bias", "decoder.conv1.weight",
            "wasp.bn1.bias", "wasp.conv1.weight", "wasp.conv2.weight", "wasp.aspp4.bn.bias", "backbone.layer4.2.bn3.bias", "backbone.layer4.2.conv3.weight",
            "backbone.layer1.2.bn3.bias", "backbone.bn1.bias", "backbone.conv1.weight"]
    print(variant, "heat", T._rel_l2(heat, ref_heat), {k: "%.1e" % T._rel_l2(params[k].grad, ref_sd[k].grad) for k in keys})
