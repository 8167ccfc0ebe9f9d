"""Drop-in for the reference's ``models.py`` on B200.

``AudioVisualGeneratorMultimodal`` keeps the reference's constructor, ``embed2out`` ModuleDict
layout (``embed2out[mod]['mu'|'log_sigma']`` are ``nn.Linear`` -- reference simplesif.py:853-856
reaches into them), ``freeze_weights`` and ``init_embedding``; its forward evaluates every head
in one libmmb_b200.so launch (``mmb_ops.HeadsFunction``) instead of 2*M cuBLAS calls + exp
kernels.  The other generator / auto-encoder classes of the reference are never instantiated
by the live script; they are kept as small PyTorch modules so that
``from models import ...`` keeps working.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import mmb_ops


def _gaussian_head(in_dim, out_dim):
    return nn.ModuleDict({'mu': nn.Linear(in_dim, out_dim), 'log_sigma': nn.Linear(in_dim, out_dim)})


def _freeze(module_dicts):
    for module in module_dicts:
        for layer in module.values():
            for param in layer.parameters():
                param.requires_grad = False


class AudioVisualGeneratorMultimodal(nn.Module):
    """reference models.py:107-202 -- MMB1 (``unimodal=True``: audio, visual) or MMB2 (audio,
    visual, audiovisual, textaudio, textvisual, textaudiovisual) Gaussian generator heads."""

    def __init__(self, embedding_dim, audio_dim, visual_dim, norm=None, frozen_weights=True,
                 unimodal=False):
        super(AudioVisualGeneratorMultimodal, self).__init__()
        self.embedding = None
        self.embedding_dim = embedding_dim

        dims = {'audio': audio_dim, 'visual': visual_dim}
        if unimodal:
            print("===========================================")
            print("Building MMB1 (unimodal factorization only)")
            print("===========================================")
        else:
            print("===========================================")
            print("Building MMB2 (uni+bi+trimodal)")
            print("===========================================")
            dims.update({
                'audiovisual': audio_dim + visual_dim,
                'textaudio': embedding_dim + audio_dim,
                'textvisual': embedding_dim + visual_dim,
                'textaudiovisual': embedding_dim + audio_dim + visual_dim,
            })
        self.embed2out = nn.ModuleDict({mod: _gaussian_head(embedding_dim, D) for mod, D in dims.items()})

        if norm is None:
            self.norm = None
        elif norm == 'layer_norm':
            self.norm = nn.LayerNorm(self.embedding_dim)
        elif norm == 'batch_norm':
            self.norm = nn.BatchNorm1d(self.embedding_dim)
        else:
            raise NotImplementedError

        if frozen_weights:
            self.freeze_weights()

    def freeze_weights(self):
        _freeze(self.embed2out.values())

    def init_embedding(self, embedding):
        assert embedding.size()[-1] == self.embedding_dim
        self.embedding = embedding
        self.embedding.requires_grad = True
        self.embedding_dim = self.embedding.size()[-1]

    def forward(self, embeddings):
        if self.norm is None:
            to_gen = embeddings
        else:
            import mmb_dp
            dp = mmb_dp.active()
            if dp is not None and isinstance(self.norm, nn.BatchNorm1d):
                # data-parallel step: the batch statistics are those of the GLOBAL batch (mmb_dp)
                to_gen = mmb_dp.sync_batch_norm(self.norm, embeddings, dp)
            else:
                to_gen = self.norm(embeddings)
        mods = list(self.embed2out.keys())
        params, is_ls = [], []
        for mod in mods:
            for name in ('mu', 'log_sigma'):
                layer = self.embed2out[mod][name]
                params.extend([layer.weight, layer.bias])
                is_ls.append(name == 'log_sigma')
        outs = mmb_ops.HeadsFunction.apply(to_gen, is_ls, *params)
        # variance must be positive: sigma = exp(log_sigma head)  (reference models.py:199)
        return {mod: {'mu': outs[2 * i], 'sigma': outs[2 * i + 1]} for i, mod in enumerate(mods)}


class AudioVisualGenerator(nn.Module):
    """reference models.py:204-253 -- the two-head predecessor (not used by the live script)."""

    def __init__(self, embedding_dim, audio_dim, visual_dim, frozen_weights=True):
        super(AudioVisualGenerator, self).__init__()
        self.embedding = None
        self.embedding_dim = embedding_dim
        self.embed2audio = _gaussian_head(embedding_dim, audio_dim)
        self.embed2visual = _gaussian_head(embedding_dim, visual_dim)
        if frozen_weights:
            self.freeze_weights()

    def freeze_weights(self):
        _freeze([self.embed2audio, self.embed2visual])

    def init_embedding(self, embedding):
        assert embedding.size()[-1] == self.embedding_dim
        self.embedding = embedding
        self.embedding.requires_grad = True
        self.embedding_dim = self.embedding.size()[-1]

    def forward(self, embeddings):
        a, v = self.embed2audio, self.embed2visual
        return ((a['mu'](embeddings), a['log_sigma'](embeddings).exp()),
                (v['mu'](embeddings), v['log_sigma'](embeddings).exp()))


class AudioVisualGeneratorConcat(nn.Module):
    """reference models.py:5-50 -- separate audio / visual latent blocks (not used)."""

    def __init__(self, audio_embedding_dim, visual_embedding_dim, audio_dim, visual_dim, frozen_weights=True):
        super(AudioVisualGeneratorConcat, self).__init__()
        self.audio_embedding_dim = audio_embedding_dim
        self.visual_embedding_dim = visual_embedding_dim
        self.embed2audio = _gaussian_head(audio_embedding_dim, audio_dim)
        self.embed2visual = _gaussian_head(visual_embedding_dim, visual_dim)

    def freeze_weights(self):
        _freeze([self.embed2audio, self.embed2visual])

    def forward(self, audio_embed, visual_embed):
        a, v = self.embed2audio, self.embed2visual
        return ((a['mu'](audio_embed), a['log_sigma'](audio_embed).exp()),
                (v['mu'](visual_embed), v['log_sigma'](visual_embed).exp()))

    def init_embeddings(self, word_embeddings):
        n = word_embeddings.size()[0]
        kw = dict(dtype=torch.float32, device=word_embeddings.device)
        return torch.cat([word_embeddings, torch.randn(n, self.audio_embedding_dim, **kw),
                          torch.randn(n, self.visual_embedding_dim, **kw)], dim=1)


class Autoencoder(nn.Module):
    """reference models.py:52-71 -- two-layer MLP auto-encoder over [text|audio|visual] (not used)."""

    def __init__(self, latent_dim, hidden_dim, embedding_dim, audio_dim, visual_dim, norm=None):
        super(Autoencoder, self).__init__()
        output_dim = embedding_dim + audio_dim + visual_dim
        self.encoder = nn.Linear(output_dim, hidden_dim)
        self.encoder2 = nn.Linear(hidden_dim, latent_dim)
        self.decoder = nn.Linear(latent_dim, hidden_dim)
        self.decoder2 = nn.Linear(hidden_dim, output_dim)

    def forward(self, inputs, device=None):
        latent = self.encoder2(F.relu(self.encoder(inputs)))
        return latent, self.decoder2(F.relu(self.decoder(latent)))


class LSTMAutoencoder(nn.Module):
    """reference models.py:73-105 -- teacher-forced LSTM sequence auto-encoder (not used)."""

    def __init__(self, latent_dim, embedding_dim, audio_dim, visual_dim):
        super(LSTMAutoencoder, self).__init__()
        output_dim = embedding_dim + audio_dim + visual_dim
        self.encoder = nn.LSTM(output_dim, latent_dim)
        self.decoder = nn.LSTM(output_dim, latent_dim)
        self.pred_layer = nn.Linear(latent_dim, output_dim)

    def forward(self, inputs, device=torch.device('cpu')):
        steps = inputs.permute(1, 0, 2)
        _, state = self.encoder(steps)
        latents = state[0]
        x = torch.zeros(1, steps.size()[1], steps.size()[2], device=device)
        preds = []
        for i in range(steps.size()[0]):
            out, state = self.decoder(x, state)
            x = steps[i:i + 1]
            preds.append(self.pred_layer(out))
        return latents, torch.cat(preds, dim=0).permute(1, 0, 2)
