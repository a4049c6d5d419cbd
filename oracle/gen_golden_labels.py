"""Golden vectors for the expert-label post-processing: the UNMODIFIED reference
``post_label_process`` (dataset/utils.py:117-160) run on small seeded uint8 label maps, with the reference's own CLIP-PCA
feature tables (dataset/*_features.pt).  Outputs + the feature rows involved go to tests/golden/prismer_labels.npz.

TEST INFRASTRUCTURE: needs /root/reference (imports ``dataset.utils`` with the reference root as working directory, because
the module loads its tables by relative path at import).

    python oracle/gen_golden_labels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PRISMER_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from tests.helpers import GOLD, label_case  # noqa: E402


def main():
    os.chdir(REF)
    sys.path.insert(0, REF)
    import dataset.utils as U
    feats = {"coco": U.COCO_FEATURES, "ade": U.ADE_FEATURES, "detection": U.DETECTION_FEATURES, "background": U.BACKGROUND_FEATURES}
    out = {"feat.coco": feats["coco"].numpy(), "feat.ade": feats["ade"].numpy(), "feat.detection": feats["detection"][:32].numpy(),
           "feat.background": feats["background"].numpy()}
    for case in range(3):
        u8, info = label_case(case)
        # what Transform.__call__ hands over (dataset/utils.py:56-63): to_tensor for depth/normal/edge, (to_tensor*255).long() else
        inputs = {}
        for k, v in u8.items():
            f = v.to(torch.float32).div(255)
            inputs[k] = f if k in ("depth", "normal", "edge") else (f * 255).long()
        res = U.post_label_process(inputs, info)
        for k, v in res.items():
            if isinstance(v, dict):
                out[f"c{case}.{k}.label"] = v["label"].numpy()
                out[f"c{case}.{k}.instance"] = v["instance"].numpy()
            else:
                out[f"c{case}.{k}"] = v.numpy()
        print(case, {k: (tuple(v["label"].shape) if isinstance(v, dict) else tuple(v.shape)) for k, v in res.items()})
    os.chdir(ROOT)
    np.savez_compressed(os.path.join(GOLD, "prismer_labels.npz"), **out)
    print(os.path.getsize(os.path.join(GOLD, "prismer_labels.npz")), "bytes")


if __name__ == "__main__":
    main()
