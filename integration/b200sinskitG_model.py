"""Reference-side binding: drop this file (and b200skitG_model.py) into the reference's `models/` directory and train with
`python train.py --model b200sinskitG ...` — `models/__init__.py:54-67` discovers a `<name>_model.py` whose class name matches
and subclasses BaseModel.  The class below IS the B200 model (`vts_b200.SinSKITGModel`, every kernel behind
include/skit_b200.h); BaseModel is mixed in only so the reference's `issubclass` test and its option plumbing accept it.

Options: the reference model's own `modify_commandline_options` (models/sinskitG_model.py:34-357) is chained first, so every
sinskitG flag keeps its name and default, then the B200-path extras (--cuda_graph, --lambda_NCE, ...) are added.
Requires the repository root (the directory holding vts_b200.py) on PYTHONPATH and the built libskit_b200.so.
"""
from models.base_model import BaseModel
from models.sinskitG_model import SinSKITGModel as _ReferenceSinSKITG

import vts_b200


class B200SinSKITGModel(vts_b200.SinSKITGModel, BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        parser = _ReferenceSinSKITG.modify_commandline_options(parser, is_train)
        # the third-party networks of the reference's default step are not on the B200 path: vision-aided (CLIP) D3 off;
        # the LPIPS-VGG16 terms need the lpips checkpoint (opt.lpips_state) and default to off here
        parser.set_defaults(use_vision_aided_loss=False, lambda_G1_lpips=0.0, lambda_G2_lpips=0.0)
        return vts_b200.SinSKITGModel.modify_commandline_options(parser, is_train)

    def __init__(self, opt):
        vts_b200.SinSKITGModel.__init__(self, opt, dist_ctx=vts_b200.dist.DistContext() if vts_b200.dist.launched_by_torchrun() else None)
