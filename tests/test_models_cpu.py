"""CPU-side checks of the model-level reference lists (tests/model_ref.py) that the GPU parity tests and bench.py's
cpu_baseline leg rely on: buffer recycling must not change a bit, zero weight zero points must be a no-op, the KV-cache
formulation must equal recomputing the prefix."""
import numpy as np

import model_ref
from rten_b200 import graphs


def _small_resnet(oracle):
    rng = oracle.XorShiftRng(5678)
    return graphs.make_resnet50(lambda s: rng.uniform(s), num_classes=10, width_mult=0.125)


def test_usable_cores(oracle):
    n = oracle.usable_cores()
    assert 1 <= n <= 4096
    assert oracle.use_all_cores() >= 1


def test_resnet_arena_is_bit_identical(oracle):
    spec = _small_resnet(oracle)
    x = oracle.XorShiftRng(1).uniform((3, 3, 64, 64))
    a = model_ref.resnet50_oracle(oracle, spec, x)
    arena = oracle.Arena()
    b = model_ref.resnet50_oracle(oracle, spec, x, arena).copy()
    c = model_ref.resnet50_oracle(oracle, spec, x, arena)  # second pass reuses every buffer
    assert np.array_equal(a.view(np.int32), b.view(np.int32)) and np.array_equal(a.view(np.int32), c.view(np.int32))


def test_int8_resnet_zero_weight_zero_points_are_a_noop(oracle):
    q = graphs.quantize_resnet50(_small_resnet(oracle))
    x = oracle.XorShiftRng(2).uniform((2, 3, 64, 64))
    la, fa = model_ref.resnet50_int8_oracle(oracle, q, x, True)
    lb, fb = model_ref.resnet50_int8_oracle(oracle, q, x, False)
    assert np.array_equal(fa.view(np.int32), fb.view(np.int32)) and np.array_equal(la.view(np.int32), lb.view(np.int32))
    assert all(np.abs(c.wq).max() <= 64 for b in q.blocks for c in (b.c1, b.c2, b.c3))  # reduce_range: 7 bits


def test_gpt2_kv_cache_steps_have_the_right_shapes_and_differ_from_scratch_only_by_quantisation(oracle):
    rng = oracle.XorShiftRng(5678)
    spec = graphs.make_gpt2_int8(lambda s: rng.uniform(s), layers=2, hidden=64, heads=4, vocab=97, max_pos=32)
    ids = (oracle.XorShiftRng(3).u64(2 * 9) % 97).astype(np.int32).reshape(2, 9)
    steps = model_ref.gpt2_int8_oracle(oracle, spec, [ids[:, :7], ids[:, 7:8], ids[:, 8:9]])
    assert [s.shape for s in steps] == [(2, 97)] * 3
    full = model_ref.gpt2_int8_oracle(oracle, spec, [ids])[0]
    # dynamic quantisation ranges differ between incremental and from-scratch runs: close, not identical
    assert np.abs(full - steps[2]).max() < 0.25 * np.abs(full).max()


def test_mnist_oracle_matches_independent_logits(oracle):
    """configs[0] (CPU plumbing): the reference's own MNIST model and the input of its own test
    (`full([1,1,28,28], 0.5)`, src/model.rs:1284-1287).  The reference pins only the output shape; the values are
    pinned here by PyTorch float64 logits stored in the fixture, to the reference's default tolerance
    (expect_equal: atol 1e-8 + rtol 1e-5 -- loosened to 1e-5 absolute for f32 accumulation over 288 terms)."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mnist.npz")
    w = graphs.load_mnist_weights(path)
    want = np.load(path)["logits_f64"]
    got = model_ref.mnist_oracle(oracle, w, np.full((1, 1, 28, 28), 0.5, np.float32))
    assert got.shape == (1, 10)
    assert np.abs(got - want).max() <= 1e-5, np.abs(got - want).max()
    assert int(got.argmax()) == int(want.argmax())


def test_generator_pairs_kv_cache_names_like_the_reference():
    """rten-generate's input/output name contract (generator.rs:283-316): `past_key_values.N.key|value` pair with
    `present.N.key|value`; a model without logits or with an unpaired cache input is rejected."""
    from rten_b200.generate import Generator

    class Fake:
        input_names = ["input_ids", "attention_mask", "past_key_values.0.key", "past_key_values.0.value", "past_key_values.1.key", "past_key_values.1.value"]
        output_names = ["logits", "present.0.key", "present.0.value", "present.1.key", "present.1.value"]

    g = Generator.from_model(Fake())
    assert g.kv_pairs == [("past_key_values.0.key", "present.0.key"), ("past_key_values.1.key", "present.1.key"),
                          ("past_key_values.0.value", "present.0.value"), ("past_key_values.1.value", "present.1.value")]
    assert g.kv_cache_len() is None

    class NoLogits(Fake):
        output_names = ["present.0.key"]

    class Unpaired(Fake):
        output_names = ["logits", "present.0.key", "present.0.value", "present.1.key"]

    import pytest
    for bad in (NoLogits, Unpaired):
        with pytest.raises(ValueError):
            Generator.from_model(bad())
