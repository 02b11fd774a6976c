"""ptbxl_multimodal_b200 -- B200-native (sm_100a) train / infer / Grad-CAM step for the
PTB-XL 1D-CNN ECG models of cyu0330/ptbxl-multimodal.  Importing the package loads
libecgb200.so; there is no CPU or library fallback."""
from ._lib import lib, EcgB200Error, EXPORTED  # noqa: F401
from .ecg_cnn import ECGCNN, ConvBlock, B200Conv1d  # noqa: F401
from .ecg_multimodal import ECGMultimodal, ECGBackbone, DemoEncoder, ECGDemoConcat  # noqa: F401
from .grad_cam_1d import GradCAM1D, gradcam_batch, compute_demo_importance  # noqa: F401
from .loop import train_one_epoch, eval_one_epoch  # noqa: F401
from .loop_demo import train_one_epoch_demo, eval_one_epoch_demo  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .infer import InferStep  # noqa: F401
from .loader import Wfdb16BatchLoader, validate_records  # noqa: F401
from . import functional, parallel, wfdb16, loader  # noqa: F401

__version__ = "0.1.0"
