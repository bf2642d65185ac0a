"""B200-native mirror of the reference's `test` package (test/__init__.py:1-22): same class names,
constructor / run() signatures, registries and on-disk contract, with the arithmetic done by
libroreg_b200.so.  See INTEGRATION.md for how Test.py picks these up unchanged."""
from .extractor import yoho_des
from .detector import yoho_det
from .matcher import mutual, yoho_mat, NMS_sample
from .estimator import yohoc, yohoo, extractor_dr_index, extractor_localtrans, yohoc_ransac, yohoo_ransac, refiner, R_pre_log

name2extractor = {'yoho_des': yoho_des}
name2detector = {'yoho_det': yoho_det}
name2matcher = {'matmul': mutual, 'yoho_mat': yoho_mat}
name2estimator = {'yohoc': yohoc, 'yohoo': yohoo}
