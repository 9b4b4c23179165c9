"""Host-side mirror of the reference's ``modules/`` package (same file names, class names,
signatures, tensor dtypes/layouts and state-dict keys) on top of the sm_100a C ABI."""
