"""Drop-in mirrors of the reference's `model` package (model/unipose.py, model/uniposeLSTM.py,
model/modules/*): same constructors, forward signatures and state_dict keys, B200 kernels underneath."""
