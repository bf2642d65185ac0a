"""Mirror of test/extractor.py:13-60 (yoho_des): FCGF_Input_Group_feature/{id}.npy -> GF network ->
YOHO_Output_Group_feature/{id}.npy (float32 [n,32,60]), skipping clouds already cached (:47)."""
import os
import numpy as np
import torch
from tqdm import tqdm
from ._common import context, make_non_exists_dir, feature_dataset_name
from .. import nets


def load_state_dict(path):
    """checkpoint['network_state_dict'] as NumPy arrays (test/extractor.py:24-27)."""
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
        self.net = nets.GFNet(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass, chunk=min(500, int(self.test_batch_size)))

    def run(self, dataset):
        self._load_model()
        datasetname = feature_dataset_name(dataset)
        FCGF_input_dir = f'{self.cfg.output_cache_fn}/{datasetname}/{self.cfg.backbone}_Input_Group_feature'
        YOHO_output_dir = f'{self.cfg.output_cache_fn}/{datasetname}/YOHO_Output_Group_feature'
        make_non_exists_dir(YOHO_output_dir)
        print(f'Extracting the PartI descriptors on {dataset.name}')
        for pc_id in tqdm(dataset.pc_ids):
            if os.path.exists(f'{YOHO_output_dir}/{pc_id}.npy'): continue
            Input_feature = np.load(f'{FCGF_input_dir}/{pc_id}.npy')          # 5000*32*60
            x = self.ctx.dev(Input_feature.astype(np.float32))
            out = self.net.forward(x)
            np.save(f'{YOHO_output_dir}/{pc_id}.npy', out.cpu().numpy())
