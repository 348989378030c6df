"""Drop-in mirror of the reference's `utils` package for the VaR hot path (B200 backend).

Same module paths and public names as the reference (`utils.factory`, `utils.calc_var_class`,
`utils.calc_var_ABC`, `utils.model_estimation.*`), new code underneath: `calc_var` runs on the GPU
through libcvar_b200.so.  Put this package's parent directory on PYTHONPATH instead of the reference's.
"""
