"""Mirror of test/detector.py:10-47 (yoho_det): rotation-guided saliency per keypoint, replaced by its rank / N
(:44-46) and written to det_score/{pc_id}.npy."""
import os
import numpy as np
from tqdm import tqdm
from ._common import context, make_non_exists_dir, feature_dataset_name
from .extractor import load_state_dict
from .. import nets


class yoho_det():
    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)
        self.best_model_fn = f'{self.cfg.model_fn}/RD/model_best.pth'
        self.npass = int(getattr(cfg, "net_passes", 3))
        self.net = nets.RDNet(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass)

    def run(self, dataset):
        datasetname = feature_dataset_name(dataset)
        savedir = f'{self.cfg.output_cache_fn}/{datasetname}/det_score'
        make_non_exists_dir(savedir)
        print(f'Evaluating the saliency of points using rotaion guided detector on {dataset.name}')
        for pc_id in tqdm(range(len(dataset.pc_ids))):
            if os.path.exists(f'{savedir}/{pc_id}.npy'): continue
            feats = np.load(f'{self.cfg.output_cache_fn}/{datasetname}/YOHO_Output_Group_feature/{pc_id}.npy')
            scores = self.net.forward(self.ctx.dev(feats.astype(np.float32))).cpu().numpy()
            # normalization for NMS comparision only (test/detector.py:44-46)
            argscores = np.argsort(scores)
            scores[argscores] = np.arange(scores.shape[0]) / scores.shape[0]
            np.save(f'{savedir}/{pc_id}.npy', scores)
