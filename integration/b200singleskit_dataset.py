"""Reference-side binding for the data pipeline: drop this file into the reference's `data/` directory and run with
`--dataset_mode b200singleskit` — `data/__init__.py:18-40` discovers `<name>_dataset.py` whose class name matches and subclasses
BaseDataset.  The class below IS the device dataset (`vts_b200.SingleSkitDataset`: Pillow-exact resize, crop, contact-centre
search and patch gathers as CUDA kernels behind include/skit_b200.h); BaseDataset is mixed in only for the `issubclass` test.

The reference wraps every dataset in `torch.utils.data.DataLoader(..., pin_memory=True)` (data/__init__.py:75-82), which cannot
take CUDA tensors, so through this binding items are handed over as host tensors (one device -> host copy per item, cached);
construct `vts_b200.SingleSkitDataset(opt)` directly to keep the items in HBM.
"""
from data.base_dataset import BaseDataset

import vts_b200


class B200SingleSkitDataset(vts_b200.SingleSkitDataset, BaseDataset):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        return vts_b200.SingleSkitDataset.modify_commandline_options(parser, is_train)

    def __init__(self, opt, verbose=False, default_len=1000):
        vts_b200.SingleSkitDataset.__init__(self, opt, verbose=verbose, default_len=default_len, host_items=True)
