import sys; sys.path.insert(0, ".")
import torch, bench
from dualpixelface_b200.synthetic import synthetic_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev, "train_faceDP", "stereodpnet").train()
bad = {}
def mk(name):
    def hook(m, inp):
        x = inp[0]
        if not x.is_contiguous(memory_format=torch.channels_last):
            bad[name] = (tuple(x.shape), x.stride())
    return hook
for n, m in model.feature_extraction.named_modules():
    if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.Conv2d)):
        m.register_forward_pre_hook(mk(n + ":" + type(m).__name__))
batch = {k: v.to(dev) for k, v in synthetic_batch(1, 256, 384, training=True, seed=0).items()}
model(batch)
for k, v in bad.items():
    print(k, v)
print(len(bad), "modules with non-channels-last input")
