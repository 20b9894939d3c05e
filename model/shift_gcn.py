"""Drop-in for the reference's model/shift_gcn.py: ``model.shift_gcn.Model`` (the dotted name the reference's
YAML configs and main.py:256-258 resolve) and the unit classes, served by the B200-native package."""
from shiftgcn_b200.modules import (Model, Shift_gcn, Shift_tcn, TCN_GCN_unit, bn_init, conv_init, import_class,  # noqa: F401
                                   tcn)
from shiftgcn_b200.shift import Shift  # noqa: F401
