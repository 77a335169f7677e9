"""Load one of the reference's own Keras checkpoints into the B200 engine and predict / continue training -- no h5py, no TensorFlow.

    python examples/load_reference_checkpoint.py /path/to/ClimSim/baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5

What the reference does with the same file: ``keras.models.load_model(f_model)`` then ``model.predict(ml_in)``
(baseline_models/MLP/training/HPO/baseline_v1/step3_prediction/step3_inference.ipynb, cell 2), or ``model.fit`` again when
``sw_continue`` is set (step2_retrain.py:252-286)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climsim_b200 import MLPEngine                                   # noqa: E402
from climsim_b200.keras_h5 import read_keras_h5                      # noqa: E402
from climsim_b200.synthetic import synthetic_batch                   # noqa: E402
from climsim_b200.trainer import Trainer, cyclical_lr                # noqa: E402


def main(path: str) -> None:
    ck = read_keras_h5(path)
    dense = ck["layers"]
    print(f"{os.path.basename(path)}: {len(dense)} Dense layers {[c['units'] for c in dense]}, "
          f"{sum(w.size for w in ck['weights'])} parameters, optimizer {ck.get('optimizer', {}).get('name')} "
          f"at iteration {ck.get('optimizer', {}).get('iterations')}")
    units = [c["units"] for c in dense[:-3]]                        # hidden widths; then Dense(128) and the [120 | 8] heads
    eng = MLPEngine.mlp_v1(units=units, dtype="bf16", max_batch=65536)
    # the file's optimizer is tfa RectifiedAdam on a cyclical learning rate: the engine's "radam" rule + trainer.cyclical_lr
    trainer = Trainer(eng, rule="radam", lr=lambda it: cyclical_lr(it, step_size=16))
    trainer.load_keras_h5(path)                                       # parameters + (m, v) + iteration count
    x, y = synthetic_batch(65536, seed=0, device="cuda")             # stand-in for normalised (N, 124) / (N, 128) columns
    pred = eng.forward(x)                                             # model.predict
    print("predictions", tuple(pred.shape), "mean |p|", float(pred.abs().mean()))
    print("one more training step of that run: loss", trainer.step(x, y))


if __name__ == "__main__":
    assert torch.cuda.is_available(), "needs a B200"
    main(sys.argv[1])
