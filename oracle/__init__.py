"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's Shift-GCN hot path (austinjeng/Shift-GCN):

* ``shift_oracle.c`` / ``shift_c.py``   scalar C restatement of the temporal-shift CUDA op (K1-K5).
* ``shift_torch.py``                    vectorised torch-CPU restatement of the same op + autograd glue.
* ``model_ref.py``                      torch-CPU restatement of Shift_gcn / Shift_tcn / TCN_GCN_unit / Model.
* ``ref_import.py`` / ``make_golden.py`` import the real reference (only in the build container, where
  /root/reference exists) to pin the restatements and to write tests/golden/*.npz.
* ``build_ref_ext.py``                  compiles the reference's own shift_cuda extension into oracle/_ref/.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (shiftgcn_b200/) never does.
"""
