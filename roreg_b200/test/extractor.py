"""Mirror of test/extractor.py:13-60 (yoho_des).  The group-conv network (SURVEY.md section 8(f)
rank 1) is a per-cloud stage whose tcgen05 implicit-GEMM kernels are not in this build."""


class yoho_des():
    def __init__(self, cfg):
        self.cfg = cfg

    def run(self, dataset):
        raise NotImplementedError("yoho_des: GF group-conv kernels are not part of this build; "
                                  "YOHO_Output_Group_feature/*.npy must be precomputed (BASELINE.json configs)")
