"""fastgen's batched replacement of HF's per-utterance post-processing (GraphedGreedy._maybe_unwind) against HF's
own code on the same finished token ids.  Runs on the CPU: the decode loop itself is replaced by a stub that returns
crafted ids, so what is compared is exactly the code after `_sample` returns -- WhisperGenerationMixin's
generate_with_fallback / _retrieve_segment / _pad_to_max_length (transformers, generation_whisper.py) on one side,
a few batched tensor ops on the other."""
import pytest
import torch

from openai_whisper_compression_b200 import fastgen, harness


@pytest.fixture(scope="module")
def rig():
    torch.manual_seed(0)
    model = harness.build_model("tiny", encoder_layers=1, decoder_layers=1).eval()
    eng = fastgen.enable(model)
    feats = torch.zeros(12, model.config.num_mel_bins, 3000)
    return model, eng, feats


def _run(model, eng, feats, crafted, fast, **kw):
    """model.generate with the decode loop stubbed to append `crafted` [B, T] to the decoder prompt."""
    def stub(input_ids, **_):       # (HF calls again with fewer rows when it cut a segment at a timestamp)
        return torch.cat([input_ids, crafted[: input_ids.shape[0]].to(input_ids.dtype)], dim=1)
    eng._orig_sample, saved = stub, eng._orig_sample
    eng.fast_post = fast
    before = eng.fast_returns
    try:
        out = model.generate(feats, do_sample=False, num_beams=1, max_new_tokens=crafted.shape[1], **kw)
    finally:
        eng._orig_sample = saved
        eng.fast_post = True
    return out, eng.fast_returns - before


def _finished_rows(B, T, eos, pad, seed, vocab=5000):
    """Rows as a greedy loop leaves them: random tokens, an eos somewhere (or nowhere), pad after it."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(3, vocab, (B, T), generator=g)
    ends = torch.randint(0, T + 4, (B,), generator=g)          # >= T: never finished
    for b in range(B):
        e = int(ends[b])
        if e < T:
            tok[b, e] = eos
            tok[b, e + 1:] = pad
    return tok


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_fast_post_equals_hf_pad_is_eos(rig, seed):
    model, eng, feats = rig
    gc = model.generation_config
    assert gc.pad_token_id == gc.eos_token_id            # Whisper's default
    tok = _finished_rows(feats.shape[0], 9 + seed, gc.eos_token_id, gc.pad_token_id, seed)
    tok[0] = torch.randint(3, 5000, (tok.shape[1],))     # a row that never finishes
    tok[1] = gc.eos_token_id                             # a row that finishes at once
    ref, n_ref = _run(model, eng, feats, tok, fast=False)
    got, n_got = _run(model, eng, feats, tok, fast=True)
    assert n_ref == 0 and n_got == 1
    assert got.dtype == ref.dtype and got.shape == ref.shape and torch.equal(got, ref)


def test_fast_post_equals_hf_all_rows_unfinished(rig):
    model, eng, feats = rig
    tok = torch.randint(3, 5000, (feats.shape[0], 16), generator=torch.Generator().manual_seed(7))
    ref, _ = _run(model, eng, feats, tok, fast=False, min_new_tokens=16)
    got, n = _run(model, eng, feats, tok, fast=True, min_new_tokens=16)
    assert n == 1 and torch.equal(got, ref) and got.shape == (feats.shape[0], 16)


def test_fast_post_equals_hf_pad_differs_from_eos(rig):
    model, eng, feats = rig
    gc = model.generation_config
    saved = gc.pad_token_id
    gc.pad_token_id = 7                                   # HF counts EVERY pad in a row that ends in pad
    try:
        tok = _finished_rows(feats.shape[0], 12, gc.eos_token_id, 7, seed=11, vocab=12)   # small vocab: stray 7s
        tok[:, 0] = 5                                     # no row is all pads (HF would index an empty row)
        ref, _ = _run(model, eng, feats, tok, fast=False)
        got, n = _run(model, eng, feats, tok, fast=True)
        assert torch.equal(got, ref)
    finally:
        gc.pad_token_id = saved


def test_fast_post_leaves_timestamp_tokens_and_non_plain_calls_to_hf(rig):
    model, eng, feats = rig
    gc = model.generation_config
    tok = _finished_rows(feats.shape[0], 10, gc.eos_token_id, gc.pad_token_id, seed=5)
    gc.no_timestamps_token_id = 4000                      # tokens >= 4001 now count as timestamps
    try:
        tok[2, 1:3] = 4500                                # two consecutive timestamp tokens: HF cuts a segment there
        ref, _ = _run(model, eng, feats, tok, fast=False)
        got, n = _run(model, eng, feats, tok, fast=True)
        assert n == 0 and torch.equal(got, ref)           # batched path declined, HF's ran
    finally:
        del gc.no_timestamps_token_id
    tok = _finished_rows(feats.shape[0], 10, gc.eos_token_id, gc.pad_token_id, seed=6)
    assert not eng._plain_call((feats,), {"return_dict_in_generate": True})
    assert not eng._plain_call((feats,), {"return_timestamps": True})
    assert not eng._plain_call((feats,), {"prompt_ids": torch.tensor([1, 2])})
    assert not eng._plain_call((feats,), {"temperature": (0.0, 0.2)})
    assert not eng._plain_call((torch.zeros(2, 80, 6000),), {})
    assert eng._plain_call((feats,), {"do_sample": False, "num_beams": 1, "max_new_tokens": 4, "language": "en"})


def test_window_preparation_shortcuts_equal_hf(rig):
    """The subclass's _maybe_reduce_batch / _get_input_segment against HF's static methods: identical results when
    the shortcut applies (first pass of a short-form batch) and when it does not (exhausted or seeked utterances,
    short windows)."""
    model, eng, feats = rig
    base, cls = eng._base_cls, type(model)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 4, 3000, generator=g)
    full = torch.full((6,), 3000, dtype=torch.long)
    cases = [(torch.zeros(6, dtype=torch.long), full.clone()),                       # first pass
             (torch.tensor([0, 3000, 0, 0, 3000, 0]), full.clone()),                 # two exhausted
             (torch.tensor([0, 0, 1200, 0, 0, 0]), full.clone())]                    # one seeked
    for seek, max_frames in cases:
        bmap = list(range(6))
        ref = base._maybe_reduce_batch(x, seek, max_frames, 6, bmap)
        got = cls._maybe_reduce_batch(x, seek, max_frames, 6, bmap)
        assert torch.equal(got[0], ref[0]) and got[1] == ref[1] and got[2] == ref[2]
        xf, bsz, bmap2 = ref
        n = (max_frames - seek).clamp(max=3000)
        ref_seg = base._get_input_segment(xf, seek, n, 3000, bsz, bmap2)
        got_seg = cls._get_input_segment(xf, seek, n, 3000, bsz, bmap2)
        assert got_seg.shape == ref_seg.shape and torch.equal(got_seg, ref_seg)
    # windows shorter than 3000 frames are padded by HF: not the shortcut's case
    short = torch.randn(3, 4, 1000, generator=g)
    z = torch.zeros(3, dtype=torch.long)
    n = torch.full((3,), 1000, dtype=torch.long)
    assert torch.equal(cls._get_input_segment(short, z, n, 3000, 3, [0, 1, 2]),
                       base._get_input_segment(short, z, n, 3000, 3, [0, 1, 2]))


def test_generate_with_all_host_shortcuts_equals_pristine_hf_on_cpu():
    """Whole `model.generate` on the CPU (HF's own decode loop): the untouched HF class against the same model after
    fastgen.enable (window shortcuts + batched post-processing active).  Same ids."""
    torch.manual_seed(1)
    model = harness.build_model("tiny", encoder_layers=1, decoder_layers=1).eval()
    feats = torch.randn(3, model.config.num_mel_bins, 3000) * 0.3
    ref = harness.greedy_generate(model, feats, 5)
    ref_free = model.generate(feats, do_sample=False, num_beams=1, max_new_tokens=6)   # eos allowed
    eng = fastgen.enable(model)
    got = harness.greedy_generate(model, feats, 5)
    got_free = model.generate(feats, do_sample=False, num_beams=1, max_new_tokens=6)
    assert eng.fast_returns == 2
    assert torch.equal(got, ref) and torch.equal(got_free, ref_free)
    eng.uninstall()
    assert torch.equal(harness.greedy_generate(model, feats, 5), ref)
