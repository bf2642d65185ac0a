"""Mirror of test/detector.py:10-47 (yoho_det) - per-cloud stage, not in this build (section 8(f) rank 4)."""


class yoho_det():
    def __init__(self, cfg):
        self.cfg = cfg

    def run(self, dataset):
        raise NotImplementedError("yoho_det: detector kernels are not part of this build")
