"""Reference-side binding for the multi-material model: `python train.py --model b200skitG ...` (see b200sinskitG_model.py).
The reference's skitG option setter (models/skitG_model.py:31-319) is chained first."""
from models.base_model import BaseModel
from models.skitG_model import SKITGModel as _ReferenceSKITG

import vts_b200


class B200SKITGModel(vts_b200.SKITGModel, BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        parser = _ReferenceSKITG.modify_commandline_options(parser, is_train)
        parser.set_defaults(use_vision_aided_loss=False, lambda_G1_lpips=0.0, lambda_G2_lpips=0.0)
        return vts_b200.SKITGModel.modify_commandline_options(parser, is_train)

    def __init__(self, opt, style_encoder=None):
        vts_b200.SKITGModel.__init__(self, opt, dist_ctx=vts_b200.dist.DistContext() if vts_b200.dist.launched_by_torchrun() else None,
                                     style_encoder=style_encoder)
