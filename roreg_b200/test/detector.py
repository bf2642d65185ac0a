"""Mirror of test/detector.py:10-47 (yoho_det): rotation-guided saliency per keypoint, replaced by its rank / N
(:44-46) and written to det_score/{pc}.npy.  As in the reference the clouds are addressed by POSITION in dataset.pc_ids
(range(len(pc_ids)), :33), and a cloud whose score file exists is skipped."""
import os
import numpy as np
from tqdm import tqdm
from ._common import context, make_non_exists_dir, CacheLayout
from .extractor import load_state_dict
from .. import nets


def rank_fraction(values):
    """test/detector.py:44-46: every value replaced by (its rank in ascending order) / N - used by the NMS comparison only."""
    out = np.asarray(values).copy()
    n = out.shape[0]
    out[np.argsort(values)] = np.arange(n) / n
    return out


class yoho_det():
    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)
        self.best_model_fn = f'{self.cfg.model_fn}/RD/model_best.pth'
        self.npass = int(getattr(cfg, "net_passes", 3))
        self.net = nets.RDNet(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass)

    def run(self, dataset):
        lay = CacheLayout(self.cfg, dataset)
        make_non_exists_dir(lay.det_score_dir)
        print(f'Evaluating the saliency of points using rotaion guided detector on {dataset.name}')
        todo = [i for i in range(len(dataset.pc_ids)) if not os.path.exists(lay.det_score(i))]
        for i in tqdm(todo):
            eqv = self.ctx.dev(np.load(lay.yoho_desc(i)).astype(np.float32))
            saliency = self.net.forward(eqv).cpu().numpy()
            np.save(lay.det_score(i), rank_fraction(saliency))
