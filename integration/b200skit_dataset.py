"""Reference-side binding for the multi-material data pipeline: drop this file into the reference's `data/` directory and run the
skitG model with `--dataset_mode b200skit` (see b200singleskit_dataset.py for how `data/__init__.py:18-40` discovers it and why items
are handed to the reference's pinning DataLoader as host tensors).  The class is `vts_b200.SkitDataset` (data/skit_dataset.py:25)."""
from data.base_dataset import BaseDataset

import vts_b200


class B200SkitDataset(vts_b200.SkitDataset, BaseDataset):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        return vts_b200.SkitDataset.modify_commandline_options(parser, is_train)

    def __init__(self, opt, verbose=False, default_len=1000):
        vts_b200.SkitDataset.__init__(self, opt, verbose=verbose, default_len=default_len, host_items=True)
