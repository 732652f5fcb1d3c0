"""TEST INFRASTRUCTURE: re-export of the parameter inventory / deterministic synthetic weights.  The definitions live
in the product package (lvd_b200/synthetic.py) because the engine and bench.py need the same inventory without
importing oracle/; they are pinned against the reference's state_dict by tests/test_oracle_pin.py."""
from lvd_b200.synthetic import (TINY, ZEROSCOPE, make_cfg, param_spec, random_state_dict,  # noqa: F401
                                up_block_resnet_channels)
