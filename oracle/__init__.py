"""CPU oracle for the ClimSim column-emulator hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch-CPU / NumPy, the arithmetic of the reference's hot path
(SURVEY.md section 8a).  It exists so that the CUDA path in ``climsim_b200`` can be checked against it.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  Nothing under ``climsim_b200/`` imports it, and the product path never falls back to it.

Parity pinning status (see DESIGN.md "Oracle"):

* ``data_utils_ref``  -- PINNED.  Checked against golden vectors produced by running the reference's own
  ``climsim_utils/data_utils.py`` (imported from /root/reference with stubbed xarray/tensorflow/netCDF4/h5py/
  matplotlib modules) on seeded synthetic arrays: ``tests/golden/make_golden.py`` -> ``tests/golden/data_utils.npz``.
* ``HSRRef``          -- PINNED.  Checked against golden vectors produced by importing the reference's own
  ``baseline_models/HSR/training/hsr.py`` (``MLP`` / ``HeteroskedasticRegression``) -> ``tests/golden/hsr_small.npz``,
  and (in this container only) against the shipped ``final_hsr.cp`` weights.
* ``MLPRef`` / ``CNNRef`` / ``EDRef`` / Keras-Adam / cyclical LR -- PARITY UNPINNED for outputs: the reference
  builds these with tensorflow 2.11.1 / keras 2.11.0 / tensorflow-addons 0.19.0 (MLP: baseline_models/MLP/env/
  environment.yml:119,287,337) and tensorflow 2.10.0 / tfa 0.18.0 (CNN, ED), none of which is installable here, and
  the reference ships no golden outputs.  They are restated from the published Keras semantics and anchored on the
  reference's call sites; the reference-held known answers they are pinned to are the parameter count
  1 753 472 (step1_results.csv, lot-147 trial_0027), the FLOP count 3 503 488 (FLOP_calculation.ipynb cells 5-6),
  the CNN/ED/HSR parameter counts derivable from the reference's layer lists, and -- STRUCTURE PINNED -- the
  ``model_config`` / ``training_config`` / weight shapes of the reference's own shipped Keras models
  (baseline_models/MLP/model/*.best.h5, baseline_models/ED/model/ED_ClimSIM_1_3_model.h5, read without h5py) and the
  layer inventory of baseline_models/CNN/model/saved_model.pb: ``tests/golden/keras_configs.json``,
  ``tests/test_oracle_pinning.py::test_graphs_and_training_setups_match_the_reference_saved_keras_models``,
  ``tests/test_keras_h5_cpu.py``.  What a TensorFlow run would OUTPUT stays unpinned.
"""
