"""Secondary metric: utterance-steps/s of ONE MMB2 latent-optimisation step
(heads -> word term + 6 masked Gaussian terms -> backward -> SGD step) at the MOSI and POM
shapes, through this repo's drop-in modules (libmmb_b200.so) against the reference's formulas
evaluated by stock PyTorch ops -- on the same B200 ("library bar") and on the host CPU cores.

    python tools/bench_mmb.py [--steps 50] [--shape mosi|pom|both] [--no-cpu]

The stock-PyTorch arm restates reference losses.py:13-34 / 68-95 and models.py:187-202 (the
CosineSimilarity broadcast over the whole vocabulary included); it is a baseline, not a checker
(parity is tests/test_mmb_gpu.py against the oracle and the golden vectors).
Prints one JSON line per shape.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import torch
import torch.nn as nn
import torch.nn.functional as F

SHAPES = {
    # name: (N, T, V, audio_dim, visual_dim)   (+2 positional columns as in make_configs.py:28)
    'mosi': (1284, 20, 3016, 74 + 2, 47 + 2),
    'pom': (600, 200, 7763, 43 + 2, 43 + 2),
    # the POM layout as the reference's main() builds it (simplesif.py:329-347, MMDataExtra): the word term
    # runs over the UNALIGNED transcript (ids padded to 1357 tokens like pom_test_ids.npy), the Gaussian terms
    # over text aligned with audio / visual (aligned length not in the tree; 200 as in SURVEY.md 8d config 3)
    'pom_real': (600, 200, 7763, 43 + 2, 43 + 2),
}
L_UNALIGNED = 1357
D, B, A_SIF = 300, 64, 1e-3


def synth(name, device, seed=0):
    N, T, V, Ad, Vd = SHAPES[name]
    g = torch.Generator(device='cpu').manual_seed(seed)
    table = 0.4 * torch.randn(V, D, generator=g) + 0.3 * torch.randn(1, D, generator=g)
    ids = torch.randint(1, V, (N, T), generator=g)
    lens = torch.randint(1, T + 1, (N, 1), generator=g)
    pad = torch.arange(T)[None, :] >= lens
    ids[pad] = 0
    text = table[ids]
    text_m = (~pad).float()[:, :, None].expand(N, T, D).contiguous()
    aud = torch.rand(N, T, Ad, generator=g) * 2 - 1
    vis = torch.rand(N, T, Vd, generator=g) * 2 - 1
    aud[pad] = -10.
    vis[pad] = -10.
    aud_m = (~pad).float()[:, :, None].expand(N, T, Ad).contiguous()
    vis_m = (~pad).float()[:, :, None].expand(N, T, Vd).contiguous()
    text_w = torch.rand(N, T, generator=g) * 0.9 + 0.1
    latents = (text * text_w[:, :, None]).sum(1) / T
    to = lambda t: t.to(device)
    out = dict(table=to(table), text=to(text), text_m=to(text_m), aud=to(aud), aud_m=to(aud_m), vis=to(vis),
               vis_m=to(vis_m), text_w=to(text_w), latents=to(latents), dims=(Ad, Vd), N=N, ids=to(ids))
    if name == 'pom_real':
        L = L_UNALIGNED
        ids_u = torch.randint(1, V, (N, L), generator=g)
        pad_u = torch.arange(L)[None, :] >= torch.randint(100, L + 1, (N, 1), generator=g)
        ids_u[pad_u] = 0
        out['text_a'], out['text_a_m'] = out['text'], out['text_m']           # aligned stream -> Gaussian terms
        tbl = out['table']
        out['text'] = tbl[ids_u.to(device)]                                      # (N, 1357, 300) -> word term
        out['text_m'] = (~pad_u).float().to(device)[:, :, None].expand(N, L, D).contiguous()
        out['text_w'] = (torch.rand(N, L, generator=g) * 0.9 + 0.1).to(device)
        out['ids_u'] = ids_u.to(device)
    return out


# ------------------------------------------------------------------ stock PyTorch restatement
def torch_gauss(mu, sigma, x, m):
    lp = torch.log(1. / torch.sqrt(2 * math.pi * sigma ** 2)) - (x - mu) ** 2 / (2 * sigma ** 2)
    return (lp * m).sum(-1).sum(-1)


def torch_word(e, table, w, sent, mask, a):
    cos = nn.CosineSimilarity(dim=-1)
    Z = (1. - torch.acos(cos(e.unsqueeze(1), table.unsqueeze(0))) / math.pi).sum(-1, keepdim=True)
    alpha = 1. / (Z * a + 1.)
    ctx = (1. - alpha) * (1. - torch.acos(cos(sent, e.unsqueeze(1))) / math.pi) / Z
    return (torch.log(alpha * w + ctx) * mask[:, :, 0]).sum(-1)


def torch_step(model, e, batch, table):
    z = model.norm(e) if model.norm is not None else e
    text, text_m, aud, aud_m, vis, vis_m, text_w = batch[:7]
    tg, tg_m = (batch[7], batch[8]) if len(batch) > 7 else (text, text_m)
    data = {'audio': aud, 'visual': vis, 'audiovisual': torch.cat([aud, vis], -1),
            'textaudio': torch.cat([tg, aud], -1), 'textvisual': torch.cat([tg, vis], -1),
            'textaudiovisual': torch.cat([tg, aud, vis], -1)}
    masks = {'audio': aud_m, 'visual': vis_m, 'audiovisual': torch.cat([aud_m, vis_m], -1),
             'textaudio': torch.cat([tg_m, aud_m], -1), 'textvisual': torch.cat([tg_m, vis_m], -1),
             'textaudiovisual': torch.cat([tg_m, aud_m, vis_m], -1)}
    total = torch_word(e, table, text_w, text, text_m, A_SIF)
    for mod, head in model.embed2out.items():
        mu = head['mu'](z).unsqueeze(1)
        sigma = head['log_sigma'](z).exp().unsqueeze(1)
        total = total + torch_gauss(mu, sigma, data[mod], masks[mod])
    return total


# ------------------------------------------------------------------ this repo's path
def ours_step(model, e, batch, word_fn, losses):
    text, text_m, aud, aud_m, vis, vis_m, text_w = batch[:7]
    tg, tg_m = (batch[7], batch[8]) if len(batch) > 7 else (text, text_m)
    C = losses.CatSegments
    data = {'text': text, 'audio': aud, 'visual': vis, 'text_weights': text_w,
            'audiovisual': C([aud, vis]), 'textaudio': C([tg, aud]), 'textvisual': C([tg, vis]),
            'textaudiovisual': C([tg, aud, vis])}
    masks = {'text': text_m, 'audio': aud_m, 'visual': vis_m, 'audiovisual': C([aud_m, vis_m]),
             'textaudio': C([tg_m, aud_m]), 'textvisual': C([tg_m, vis_m]),
             'textaudiovisual': C([tg_m, aud_m, vis_m])}
    out = model(e)
    return losses.get_log_prob_matrix({}, e, out, data, masks, word_fn, device=e.device)


def run(name, steps, do_cpu, only_ours=False):
    import losses
    import models
    import simplesif
    dev = torch.device('cuda')
    S = synth(name, dev)
    Ad, Vd = S['dims']
    torch.manual_seed(0)
    model = models.AudioVisualGeneratorMultimodal(D, Ad, Vd, norm='layer_norm', frozen_weights=False).to(dev)
    word_fn = simplesif.make_word_log_prob_fn({'word_sim_metric': 'angular'}, None, S['table'], a=A_SIF)
    n_batches = S['N'] // B
    perm = torch.randperm(S['N'], generator=torch.Generator().manual_seed(1)).to(dev)

    def make_loop(step_fn, device):
        Sd = S if device.type == 'cuda' else {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in S.items()}
        mdl = model if device.type == 'cuda' else \
            models.AudioVisualGeneratorMultimodal(D, Ad, Vd, norm='layer_norm', frozen_weights=False)
        lat = Sd['latents'].clone().requires_grad_(True)
        opt = torch.optim.SGD([lat] + list(mdl.parameters()), lr=1e-7)  # tiny: synthetic data, timing only
        pm = perm.to(device)

        def one(i):
            j = pm[(i % n_batches) * B:(i % n_batches + 1) * B]
            keys = ('text', 'text_m', 'aud', 'aud_m', 'vis', 'vis_m', 'text_w') + (
                ('text_a', 'text_a_m') if 'text_a' in Sd else ())
            batch = tuple(Sd[k][j] for k in keys)
            opt.zero_grad()
            lp = step_fn(mdl, lat[j], batch)
            loss = (-lp).mean()
            loss.backward()
            opt.step()
            return loss
        return one

    def time_gpu(one, n):
        for i in range(5):
            one(i)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n):
            loss = one(i)
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n, float(loss.detach())

    ours = make_loop(lambda m, e, b: ours_step(m, e, b, word_fn, losses), dev)
    ms_ours, l_ours = time_gpu(ours, steps)
    if only_ours:
        print(json.dumps({'shape': name, 'b200_fused': {'ms_per_step': ms_ours}}))
        return
    # the same step replayed as one captured CUDA graph (simplesif.GraphedStep)
    import utils
    mk = {'text': S['text_m'], 'covarep': S['aud_m'], 'facet': S['vis_m']}
    if 'text_a' in S:
        ds = utils.MMDataExtra(S['text'], S['aud'], S['vis'], dict(mk, text_align=S['text_a_m']), S['text_w'],
                               S['text_a'], dev)
    else:
        ds = utils.MMData(S['text'], S['aud'], S['vis'], mk, S['text_w'], dev)
    lat_g = S['latents'].clone().requires_grad_(True)
    opt_g = torch.optim.SGD([lat_g] + list(model.parameters()), lr=1e-7)
    stepper = simplesif.GraphedStep({'dataset': 'pom' if 'text_a' in S else 'mosi', 'unimodal': False}, model, lat_g,
                                    ds, opt_g, word_fn, dev)
    ms_graph, l_graph = time_gpu(lambda i: stepper(perm[(i % n_batches) * B:(i % n_batches + 1) * B]), steps)
    stepper.check()
    ids_arm = None
    if True:
        # the transcript as ids (no (N, L, 300) tensors), word term from the ids alone
        ones = torch.ones(S['table'].shape[0], device=dev)
        if 'ids_u' in S:
            ds_i = utils.MMDataExtraIds(S['ids_u'], S['aud'], S['vis'], dict(mk, text_align=S['text_a_m']), ones,
                                        S['table'], S['text_a'], dev)
        else:
            ds_i = utils.MMDataIds(S['ids'], S['aud'], S['vis'], mk, ones, S['table'], dev)
        ds_i.text_weights = S['text_w']
        lat_i = S['latents'].clone().requires_grad_(True)
        opt_i = torch.optim.SGD([lat_i] + list(model.parameters()), lr=1e-7)
        st_i = simplesif.GraphedStep({'dataset': 'pom' if 'ids_u' in S else 'mosi', 'unimodal': False}, model, lat_i,
                                     ds_i, opt_i, word_fn, dev)
        ms_i, l_i = time_gpu(lambda i: st_i(perm[(i % n_batches) * B:(i % n_batches + 1) * B]), steps)
        st_i.check()
        ids_arm = {'ms_per_step': ms_i, 'value': B / ms_i * 1e3, 'loss': l_i}
    stock = make_loop(lambda m, e, b: torch_step(m, e, b, S['table']), dev)
    ms_stock, l_stock = time_gpu(stock, max(5, steps // 5))
    res = {'metric': 'utterance-steps/s (one MMB2 fwd+bwd+SGD step, B=64)', 'shape': name,
           'N_T_V_A_Vd': SHAPES[name], 'steps': steps,
           'b200_fused': {'ms_per_step': ms_ours, 'value': B / ms_ours * 1e3, 'loss': l_ours},
           'b200_fused_cuda_graph': {'ms_per_step': ms_graph, 'value': B / ms_graph * 1e3, 'loss': l_graph},
           'b200_stock_torch': {'ms_per_step': ms_stock, 'value': B / ms_stock * 1e3, 'loss': l_stock}}
    if ids_arm:
        res['b200_fused_cuda_graph_token_ids'] = ids_arm
    if do_cpu:
        cpu = make_loop(lambda m, e, b: torch_step(m, e, b, S['table'].cpu()), torch.device('cpu'))
        cpu(0)
        t0 = time.perf_counter()
        n = 3
        for i in range(n):
            cpu(i)
        dt = (time.perf_counter() - t0) / n
        res['cpu_stock_torch'] = {'ms_per_step': dt * 1e3, 'value': B / dt, 'cores': torch.get_num_threads()}
    print(json.dumps(res))
    return res


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--shape', default='both')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--only-eager-ours', action='store_true', help='profiling aid: only the eager fused arm')
    a = ap.parse_args()
    for nm in (['mosi', 'pom'] if a.shape == 'both' else ['mosi', 'pom', 'pom_real'] if a.shape == 'all' else [a.shape]):
        run(nm, a.steps, not a.no_cpu, a.only_eager_ours)
