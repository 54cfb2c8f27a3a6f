"""`--framework b200-keras`: the engine with the Keras backend's numerics and schedule
(keras_model.py, keras_topk_word_predictions_layer.py,
keras_words_subtoken_metrics.py).  Checked: full-vocabulary softmax scores against the oracle,
the Keras initialisers' ranges, Adam with epsilon 1e-7 against the oracle, and train / evaluate /
predict / resume through the model surface."""
import os

import numpy as np
import pytest

from oracle import path_attention_oracle as O
from tests.test_gpu_model import _config, _make_dataset
from tests.util import dev_batch, make_engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,B", [
    (O.Dims(token_vocab=1001, path_vocab=501, target_vocab=1001, embed_dim=32, code_dim=96, max_contexts=20), 64),
    (O.Dims(token_vocab=300, path_vocab=200, target_vocab=1537, embed_dim=20, code_dim=52, max_contexts=13), 7),
])
def test_full_vocabulary_softmax_scores(dims, B):
    params = O.keras_init_params(dims, seed=11)
    eng, _ = make_engine(dims, max_batch=B, training=False, params=params)
    src, pth, tgt, mask, _ = O.synthetic_batch(dims, B, seed=3)
    idx, probs, _, _ = eng.predict_batch_host(src, pth, tgt, mask, normalize=2)
    oi, op, _, _, scores = O.evaluate_topk(params, src, pth, tgt, mask, k=10, normalize=2)
    assert np.array_equal(idx, oi)
    assert np.abs(probs - op).max() < 2e-6 * max(1.0, float(op.max()))
    # they are probabilities of the whole vocabulary: descending, < 1 in total, and the k-softmax differs
    assert np.all(np.diff(probs, axis=1) <= 0) and np.all(probs.sum(axis=1) < 1.0)
    full = np.exp(scores - scores.max(1, keepdims=True))
    full /= full.sum(1, keepdims=True)
    assert np.abs(probs - np.take_along_axis(full, idx, axis=1)).max() < 1e-6
    # device entry point agrees with the host one
    import torch
    code, _ = eng.forward(*dev_batch(eng, src, pth, tgt, mask))
    di, dv = eng.topk(code, normalize=2)
    torch.cuda.synchronize()
    assert np.array_equal(di.cpu().numpy(), idx) and np.allclose(dv.cpu().numpy(), probs, atol=1e-7)


def test_keras_initialisers_and_adam_epsilon():
    dims = O.Dims(token_vocab=2001, path_vocab=1001, target_vocab=801, embed_dim=32, code_dim=96, max_contexts=10)
    eng, _ = make_engine(dims, max_batch=8)
    eng.init_params(seed=5, scheme="keras")
    p = eng.export_params()
    lim = {"tok": 0.05, "path": 0.05, "a": 0.05, "W": np.sqrt(6.0 / (96 + 96)), "tgt": np.sqrt(6.0 / (96 + 801))}
    for k, l in lim.items():
        assert np.abs(p[k]).max() <= l + 1e-7 and np.abs(p[k]).max() > 0.9 * l, k
        assert abs(float(p[k].mean())) < 0.05 * l, k
    with pytest.raises(ValueError):
        eng.init_params(scheme="nope")
    # three Adam steps with the Keras epsilon against the oracle's TF-faithful Adam
    params = O.keras_init_params(dims, seed=7)
    eng.load_params(params)
    ref = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in ref.items()}
    v = {k: np.zeros_like(x) for k, x in ref.items()}
    for s in range(3):
        src, pth, tgt, mask, target = O.synthetic_batch(dims, 8, seed=60 + s)
        eng.train_step(*dev_batch(eng, src, pth, tgt, mask, target), keep=1.0)
        eng.adam_step(eps=O.KERAS_ADAM_EPS)
        _, g, _ = O.train_loss_and_grads(ref, src, pth, tgt, mask, target)
        O.adam_step(ref, g, m, v, s + 1, eps=O.KERAS_ADAM_EPS)
    got = eng.export_params()
    for k in O.PARAM_NAMES:
        assert np.abs(got[k] - ref[k]).max() < 5e-5, k


def test_keras_backend_train_evaluate_predict_resume(tmp_path, monkeypatch):
    from code2vec_b200 import load_model_dynamically
    from code2vec_b200.b200_keras_model import Code2VecModel as KerasNumericsModel
    monkeypatch.chdir(tmp_path)
    prefix, test_lines = _make_dataset(tmp_path)
    save_path = str(tmp_path / "kmodel" / "saved")
    cfg = _config(prefix, tmp_path, TRAIN_DATA_PATH_PREFIX=prefix, MODEL_SAVE_PATH=save_path, TEST_DATA_PATH=prefix + ".test.c2v",
                  DROPOUT_KEEP_RATE=1.0, DL_FRAMEWORK="b200-keras", NUM_TRAIN_EPOCHS=150, SAVE_EVERY_EPOCHS=150,
                  NUM_TRAIN_BATCHES_TO_EVALUATE=10 ** 6)
    logged = []
    model = load_model_dynamically(cfg)
    assert isinstance(model, KerasNumericsModel) and model.trainer.adam["eps"] == 1e-7
    monkeypatch.setattr(model, "log", lambda msg: logged.append(str(msg)))
    evals = []
    real_eval = model.evaluate
    monkeypatch.setattr(model, "evaluate", lambda: evals.append(real_eval()) or evals[-1])
    model.train()
    assert model.nr_epochs_trained == 150 and len(evals) == 150            # once per epoch end (ModelEvaluationCallback)
    assert os.path.exists(save_path + ".c2v_b200")
    assert any(s.startswith("Completed epoch #150") for s in logged) and any("top1:" in s for s in logged)
    res = evals[-1]
    assert len(res.topk_acc) == 5 and res.topk_acc[0] > 0.6 and np.all(np.diff(res.topk_acc) >= 0)
    assert res.loss is not None and res.loss < evals[0].loss and 0.0 < res.subtoken_f1 <= 1.0
    trained = model.engine.export_params()

    # the evaluation loss is the mean cross entropy of the test set under the exported parameters
    from code2vec_b200.path_context_reader import EstimatorAction, PathContextReader
    from code2vec_b200.b200_keras_model import _KerasEvaluateInputFormer
    reader = PathContextReader(vocabs=model.vocabs, model_input_tensors_former=_KerasEvaluateInputFormer(), config=cfg,
                               estimator_action=EstimatorAction.Evaluate)
    tot, n = 0.0, 0
    for batch in reader.get_dataset():
        src, pth, tgt, mask, target, _ = batch
        v_, _, _ = O.forward(trained, src, pth, tgt, mask)
        loss, per, _ = O.softmax_xent(O.logits_of(trained, v_), np.asarray(target).reshape(-1))
        tot += float(np.sum(per))
        n += len(per)
    assert n == 24 and abs(tot / n - res.loss) < 1e-4

    preds = model.predict(test_lines[:2])
    p = preds[0]
    assert len(p.topk_predicted_words) == 5 and 0.0 < float(np.sum(p.topk_predicted_words_scores)) <= 1.0 + 1e-6
    assert p.code_vector is not None and p.code_vector.shape == (48,)
    model.close_session()

    # resume: the checkpoint carries the number of epochs trained, so fit() has nothing left to do
    cfg2 = _config(prefix, tmp_path, TRAIN_DATA_PATH_PREFIX=prefix, MODEL_LOAD_PATH=save_path, TEST_DATA_PATH=prefix + ".test.c2v",
                   DL_FRAMEWORK="b200-keras", NUM_TRAIN_EPOCHS=150)
    m2 = load_model_dynamically(cfg2)
    assert m2.nr_epochs_trained == 150
    for k, arr in m2.engine.export_params().items():
        assert np.array_equal(arr, trained[k]), k
    m2.close_session()
