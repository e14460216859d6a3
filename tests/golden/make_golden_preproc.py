"""Generates tests/golden/golden_preproc.pt by running THE REFERENCE's SegDataPreProcessor
(mmseg/models/data_preprocessor.py, executed through oracle/ref_loader.py) in the build container.
Usage:  python tests/golden/make_golden_preproc.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


class _DS:
    def set_metainfo(self, d):
        self.meta = d


def main():
    cls = ref_loader.load_data_preprocessor()
    g = torch.Generator().manual_seed(99)
    cases = []
    # (H, W, kwargs of the constructor): the config's own setting (cfg:13-20) first
    for H, W, kw in [(16, 24, dict(mean=MEAN, std=STD, bgr_to_rgb=True, size=(16, 24), pad_val=0, seg_pad_val=255)),
                     (13, 9, dict(mean=MEAN, std=STD, bgr_to_rgb=True, pad_val=0, test_cfg=dict(size=(16, 12)))),
                     (10, 10, dict(mean=MEAN, std=STD, rgb_to_bgr=True, pad_val=1.5, test_cfg=dict(size_divisor=8))),
                     (7, 5, dict(mean=MEAN, std=STD)),
                     (6, 6, dict(bgr_to_rgb=True))]:
        imgs = [torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8) for _ in range(3)]
        imgs[0][:, 0, 0] = torch.tensor([0, 255, 128], dtype=torch.uint8)        # extremes
        pp = cls(**kw)
        out = pp(dict(inputs=[i.clone() for i in imgs], data_samples=[_DS() for _ in imgs]), training=False)["inputs"]
        cases.append(dict(H=H, W=W, kw=kw, imgs=torch.stack(imgs), out=out.clone()))
    # every byte value through the config's normalisation: the full 3 x 256 table
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 16, 16).expand(3, 16, 16).contiguous()
    pp = cls(mean=MEAN, std=STD, bgr_to_rgb=True, size=(16, 16))
    table = pp(dict(inputs=[ramp]), training=False)["inputs"]
    torch.save(dict(cases=cases, ramp=ramp, table=table), os.path.join(HERE, "golden_preproc.pt"))
    print("wrote golden_preproc.pt:", len(cases), "cases;", [tuple(c["out"].shape) for c in cases])


if __name__ == "__main__":
    main()
