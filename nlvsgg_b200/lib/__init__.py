"""Drop-in mirrors of the reference's lib/ modules on the hot path (same names, constructors, forward(entry)
contract and state_dict keys), backed by the sm_100a kernels.  Put this package's parent on sys.path ahead of the
reference tree — see INTEGRATION.md."""
