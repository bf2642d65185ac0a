"""Mirror of test/extractor.py:13-60 (yoho_des): FCGF_Input_Group_feature/{id}.npy -> GF network ->
YOHO_Output_Group_feature/{id}.npy (float32 [n,32,60]); a cloud whose output file exists is skipped (:47)."""
import os
import numpy as np
import torch
from tqdm import tqdm
from ._common import context, make_non_exists_dir, CacheLayout
from .. import nets


def load_state_dict(path):
    """checkpoint['network_state_dict'] as NumPy arrays (test/extractor.py:24-27); a missing file is the reference's
    ValueError("No model exists")."""
    if not os.path.exists(path):
        raise ValueError("No model exists")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    return {k: v.detach().cpu().numpy() for k, v in ck["network_state_dict"].items()}


class yoho_des():
    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)
        self.best_model_fn = f'{self.cfg.model_fn}/GF/model_best.pth'
        self.test_batch_size = self.cfg.bs_GF
        self.npass = int(getattr(cfg, "net_passes", 3))
        self.net = None

    def _load_model(self):
        self.net = nets.GFNet(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass, chunk=max(5000, int(self.test_batch_size)))   # bs_GF is the reference's memory knob; the output does not depend on the chunking (eval-mode BN)

    def run(self, dataset):
        self._load_model()
        lay = CacheLayout(self.cfg, dataset)
        make_non_exists_dir(lay.yoho_dir)
        print(f'Extracting the PartI descriptors on {dataset.name}')
        todo = [pc for pc in dataset.pc_ids if not os.path.exists(lay.yoho_desc(pc))]
        for pc in tqdm(todo):
            fcgf = self.ctx.dev(np.load(lay.fcgf_desc(pc)).astype(np.float32))      # [n,32,60]
            np.save(lay.yoho_desc(pc), self.net.forward(fcgf).cpu().numpy())
