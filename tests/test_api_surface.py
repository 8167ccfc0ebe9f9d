"""CPU: the drop-in boundary of SURVEY.md 8(b) -- every function / method of the reference's module-level API on the
hot path exists here under the same name with the same leading parameters in the same order (callers pass
positionally, e.g. reference sif.py:82, 92; simplesif.py:129-131, 588-590) and at least as many of them optional.
Parameters this repo adds must come after the reference's and be optional.  The table below was extracted from the
reference's sources with `ast` (names and defaults only -- no code)."""
import importlib
import inspect

import pytest

REFERENCE_SIGNATURES = {
 "losses:full_loss": {
  "args": [
   "predictions",
   "y_test"
  ],
  "n_defaults": 0
 },
 "losses:get_log_prob_matrix": {
  "args": [
   "args",
   "latents",
   "out",
   "data",
   "masks",
   "word_log_prob_fn",
   "device",
   "verbose"
  ],
  "n_defaults": 2
 },
 "losses:get_log_prob_matrix_old": {
  "args": [
   "args",
   "latents",
   "audio",
   "visual",
   "data",
   "masks",
   "word_log_prob_fn",
   "device",
   "verbose"
  ],
  "n_defaults": 2
 },
 "losses:get_normal_log_prob": {
  "args": [
   "mu",
   "sigma",
   "values",
   "mask"
  ],
  "n_defaults": 0
 },
 "losses:get_word_log_prob_angular": {
  "args": [
   "latents",
   "weights",
   "word_embeddings",
   "data",
   "mask",
   "a"
  ],
  "n_defaults": 0
 },
 "losses:get_word_log_prob_angular2": {
  "args": [
   "latents",
   "word_embeddings",
   "word_weights",
   "sent_embeddings",
   "mask",
   "a"
  ],
  "n_defaults": 0
 },
 "losses:get_word_log_prob_dot_prod": {
  "args": [
   "latents",
   "weights",
   "word_embeddings",
   "data",
   "a"
  ],
  "n_defaults": 0
 },
 "losses:get_word_log_prob_dot_prod2": {
  "args": [
   "latents",
   "word_embeddings",
   "word_weights",
   "sent_embeddings",
   "mask",
   "a"
  ],
  "n_defaults": 0
 },
 "losses:iemocap_loss": {
  "args": [
   "predictions",
   "y_test"
  ],
  "n_defaults": 0
 },
 "losses:pom_loss": {
  "args": [
   "predictions",
   "y_test"
  ],
  "n_defaults": 0
 },
 "models:AudioVisualGeneratorMultimodal.__init__": {
  "args": [
   "self",
   "embedding_dim",
   "audio_dim",
   "visual_dim",
   "norm",
   "frozen_weights",
   "unimodal"
  ],
  "n_defaults": 3
 },
 "models:AudioVisualGeneratorMultimodal.forward": {
  "args": [
   "self",
   "embeddings"
  ],
  "n_defaults": 0
 },
 "models:AudioVisualGeneratorMultimodal.freeze_weights": {
  "args": [
   "self"
  ],
  "n_defaults": 0
 },
 "models:AudioVisualGeneratorMultimodal.init_embedding": {
  "args": [
   "self",
   "embedding"
  ],
  "n_defaults": 0
 },
 "sentiment_model:SentimentData.__getitem__": {
  "args": [
   "self",
   "idx"
  ],
  "n_defaults": 0
 },
 "sentiment_model:SentimentData.__init__": {
  "args": [
   "self",
   "sentiment",
   "device"
  ],
  "n_defaults": 0
 },
 "sentiment_model:SentimentData.__len__": {
  "args": [
   "self"
  ],
  "n_defaults": 0
 },
 "sentiment_model:SentimentModel.__init__": {
  "args": [
   "self",
   "embedding_dim",
   "hidden_dim",
   "n_out"
  ],
  "n_defaults": 0
 },
 "sentiment_model:SentimentModel.forward": {
  "args": [
   "self",
   "inputs"
  ],
  "n_defaults": 0
 },
 "sentiment_model:predict_sentiment": {
  "args": [
   "data",
   "model",
   "latents"
  ],
  "n_defaults": 0
 },
 "sentiment_model:save_sentiment": {
  "args": [
   "path",
   "model"
  ],
  "n_defaults": 0
 },
 "sentiment_model:train_sentiment": {
  "args": [
   "args",
   "model",
   "train_data",
   "train_latents",
   "valid_data",
   "valid_latents",
   "model_loader",
   "valid_niter",
   "verbose",
   "model_save_path"
  ],
  "n_defaults": 3
 },
 "sentiment_model:train_sentiment_for_latents": {
  "args": [
   "args",
   "latents",
   "sentiment_data",
   "device",
   "verbose",
   "model_save_path",
   "train_idxes"
  ],
  "n_defaults": 3
 },
 "sif2:calc_weights": {
  "args": [
   "data",
   "b_mean",
   "b_log_sigma",
   "mask"
  ],
  "n_defaults": 0
 },
 "sif2:estimate_embedding_overall_gpu2": {
  "args": [
   "data",
   "masks",
   "networks",
   "sentence_weights",
   "embeddings"
  ],
  "n_defaults": 0
 },
 "sif:get_sentence_embeddings": {
  "args": [
   "word_embeddings",
   "weights",
   "text"
  ],
  "n_defaults": 0
 },
 "sif:get_sentence_word_weights": {
  "args": [
   "text",
   "weights"
  ],
  "n_defaults": 0
 },
 "sif:get_word_weights": {
  "args": [
   "word_freq_file",
   "a"
  ],
  "n_defaults": 1
 },
 "sif:load_iemocap_weights": {
  "args": [],
  "n_defaults": 0
 },
 "sif:load_mosi_weights": {
  "args": [],
  "n_defaults": 0
 },
 "sif:load_pom_weights": {
  "args": [],
  "n_defaults": 0
 },
 "sif:load_weights": {
  "args": [
   "args"
  ],
  "n_defaults": 0
 },
 "sif_functions:SIF_embedding": {
  "args": [
   "We",
   "x",
   "w",
   "params"
  ],
  "n_defaults": 0
 },
 "sif_functions:compute_pc": {
  "args": [
   "X",
   "npc"
  ],
  "n_defaults": 1
 },
 "sif_functions:get_weighted_average": {
  "args": [
   "We",
   "x",
   "w"
  ],
  "n_defaults": 0
 },
 "sif_functions:remove_pc": {
  "args": [
   "X",
   "npc"
  ],
  "n_defaults": 1
 },
 "sif_functions:seq2weight": {
  "args": [
   "seq",
   "mask",
   "weight4ind"
  ],
  "n_defaults": 0
 },
 "simplesif:optimize_latents": {
  "args": [
   "args",
   "train",
   "gen_model",
   "embed_arr",
   "dataloader",
   "n_epochs",
   "lr",
   "word_prob_fn",
   "device",
   "validation_data",
   "verbose"
  ],
  "n_defaults": 2
 },
 "simplesif:parse_arguments": {
  "args": [],
  "n_defaults": 0
 },
 "simplesif:read_config": {
  "args": [
   "config_file"
  ],
  "n_defaults": 0
 },
 "simplesif:update_masks": {
  "args": [
   "mask_dict",
   "data",
   "embedding_dim"
  ],
  "n_defaults": 0
 },
 "simplesif:update_masks_vect": {
  "args": [
   "mask_dict",
   "data",
   "key"
  ],
  "n_defaults": 1
 },
 "utils:MMData.__getitem__": {
  "args": [
   "self",
   "idx"
  ],
  "n_defaults": 0
 },
 "utils:MMData.__init__": {
  "args": [
   "self",
   "text",
   "audio",
   "visual",
   "masks",
   "text_weights",
   "device"
  ],
  "n_defaults": 0
 },
 "utils:MMData.__len__": {
  "args": [
   "self"
  ],
  "n_defaults": 0
 },
 "utils:MMDataExtra.__getitem__": {
  "args": [
   "self",
   "idx"
  ],
  "n_defaults": 0
 },
 "utils:MMDataExtra.__init__": {
  "args": [
   "self",
   "text",
   "audio",
   "visual",
   "masks",
   "text_weights",
   "text_aligned",
   "device"
  ],
  "n_defaults": 0
 },
 "utils:add_positional_embeddings": {
  "args": [
   "args",
   "data"
  ],
  "n_defaults": 0
 },
 "utils:load_data": {
  "args": [
   "args"
  ],
  "n_defaults": 0
 },
 "utils:load_iemocap": {
  "args": [
   "args"
  ],
  "n_defaults": 0
 },
 "utils:load_mosi": {
  "args": [],
  "n_defaults": 0
 },
 "utils:load_pom": {
  "args": [],
  "n_defaults": 0
 },
 "utils:normalize_data": {
  "args": [
   "train"
  ],
  "n_defaults": 0
 }
}


def _resolve(key):
    mod_name, qual = key.split(':')
    obj = importlib.import_module(mod_name)
    for part in qual.split('.'):
        obj = getattr(obj, part)
    return obj


@pytest.mark.parametrize('key', sorted(REFERENCE_SIGNATURES))
def test_reference_call_surface(key):
    want = REFERENCE_SIGNATURES[key]
    fn = _resolve(key)
    params = list(inspect.signature(fn).parameters.values())
    if want['args'] and want['args'][0] == 'self' and (not params or params[0].name != 'self'):
        want = dict(want, args=want['args'][1:])          # bound through the class: inspect drops nothing, AST keeps self
    names = [p.name for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    n_ref = len(want['args'])
    assert names[:n_ref] == want['args'], (key, names, want['args'])
    n_required_ref = n_ref - want['n_defaults']
    for i, p in enumerate(params[:len(names)]):
        if i >= n_required_ref:
            assert p.default is not inspect.Parameter.empty, (key, p.name, 'must be optional: the reference gives it a default'
                                                              if i < n_ref else 'extra parameters must be optional')
