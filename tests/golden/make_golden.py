"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``python tests/golden/make_golden.py``): it imports
``sif_functions`` / ``sif`` / ``losses`` / ``models`` from /root/reference, feeds them the
seeded inputs built by ``tests/golden/cases.py`` and stores inputs that are small (or the
seed that regenerates them, plus a checksum) together with the reference's outputs.
/root/reference does not exist on the GPU box; tests only read the committed ``.npz``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MMB_REFERENCE', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, HERE)

import torch  # noqa: E402

import sif_functions as ref_sf  # noqa: E402  (reference)
import sif as ref_sif  # noqa: E402  (reference)
import losses as ref_losses  # noqa: E402  (reference)
import models as ref_models  # noqa: E402  (reference)

import cases  # noqa: E402  (ours: seeded input builders shared with the tests)

assert ref_sf.__file__.startswith(REF), ref_sf.__file__


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print('wrote', name, {k: getattr(v, 'shape', None) for k, v in arrs.items()},
          os.path.getsize(path) // 1024, 'KiB')


def sif_case(name, We, weights, ids, store_inputs):
    w = ref_sf.seq2weight(ids, np.ones(ids.shape), weights)
    avg = ref_sf.get_weighted_average(We, ids, w)
    pc = ref_sf.compute_pc(avg, 1)
    emb = ref_sif.get_sentence_embeddings(We, weights, ids)
    keep = slice(0, 64)   # the fixtures stay small: first 64 rows + whole-array checksums
    out = dict(w=w, avg=avg[keep], pc=pc, emb=emb[keep], avg_sum=cases.checksum(avg),
               emb_sum=cases.checksum(emb), table_sum=cases.checksum(We))
    if store_inputs:
        out.update(We=We, weights=weights, ids=ids)
    save(name, **out)


def main():
    # ---- SIF: MOSI-like small case, inputs stored (N < 300 -> sklearn transposes) -----
    We, weights, ids = cases.sif_mosi_like()
    sif_case('sif_mosi_like.npz', We, weights, ids, store_inputs=True)

    # ---- SIF: N >= 300 rows, short rows, negative ids present ---------------------------
    We, weights, ids = cases.sif_tall()
    sif_case('sif_tall.npz', We, weights, ids, store_inputs=True)

    # ---- SIF: real POM test ids (first 8 utterances, L=1357) + real POM word weights ----
    pom_ids = np.load(os.path.join(REF, 'pom', 'pom_test_ids.npy'))
    pom_w = np.load(os.path.join(REF, 'pom', 'pom_word_weights.npy')).squeeze()
    ids = pom_ids[:8]
    We = cases.table(7763, 300, seed=11)
    w = ref_sf.seq2weight(ids, np.ones(ids.shape), pom_w)
    avg = ref_sf.get_weighted_average(We, ids, w)
    emb = ref_sif.get_sentence_embeddings(We, pom_w, ids)
    save('sif_pom_real.npz', ids=ids.astype(np.int16), weights=pom_w, w_rowsum=w.sum(1),
         w_nonzero=np.count_nonzero(w, axis=1), avg=avg, emb=emb, pc=ref_sf.compute_pc(avg, 1),
         table_sum=cases.checksum(We),
         valid_shape=np.array(np.load(os.path.join(REF, 'pom', 'pom_valid_ids.npy')).shape))

    # ---- compute_pc / remove_pc on bare matrices, npc in {1,2,3}, both N regimes --------
    out = {}
    for tag, (n, gap, seed) in cases.PC_CASES.items():
        X = cases.pc_matrix(n, gap, seed)
        out[tag + '_sum'] = cases.checksum(X)
        for npc in (1, 2, 3):
            out['%s_pc%d' % (tag, npc)] = ref_sf.compute_pc(X, npc)
            out['%s_rm%d' % (tag, npc)] = ref_sf.remove_pc(X, npc)[:16]
    save('pc_cases.npz', **out)

    # ---- MMB: losses + heads, values and autograd gradients -----------------------------
    for tag, cfg in cases.MMB_CASES.items():
        c = cases.mmb_inputs(**cfg)
        torch.manual_seed(cfg['seed'])
        model = ref_models.AudioVisualGeneratorMultimodal(
            c['d'], c['A'], c['Vd'], norm=cfg['norm'], frozen_weights=False,
            unimodal=cfg['unimodal'])
        cases.load_heads(model, c['heads'], c.get('norm_params'))
        lat = torch.tensor(c['latents'], requires_grad=True)
        out_heads = model(lat)
        t = {k: torch.tensor(v) for k, v in c.items() if isinstance(v, np.ndarray)}
        text_gauss, text_gauss_m = t['text'], t['text_m']
        if cfg['unimodal']:
            data = {'text': t['text'], 'audio': t['aud'], 'visual': t['vis'],
                    'text_weights': t['text_w']}
            masks = {'text': t['text_m'], 'audio': t['aud_m'], 'visual': t['vis_m']}
        else:  # simplesif.py:94-113
            data = {'text': t['text'], 'audio': t['aud'], 'visual': t['vis'],
                    'text_weights': t['text_w'],
                    'audiovisual': torch.cat([t['aud'], t['vis']], -1),
                    'textaudio': torch.cat([text_gauss, t['aud']], -1),
                    'textvisual': torch.cat([text_gauss, t['vis']], -1),
                    'textaudiovisual': torch.cat([text_gauss, t['aud'], t['vis']], -1)}
            masks = {'text': t['text_m'], 'audio': t['aud_m'], 'visual': t['vis_m'],
                     'audiovisual': torch.cat([t['aud_m'], t['vis_m']], -1),
                     'textaudio': torch.cat([text_gauss_m, t['aud_m']], -1),
                     'textvisual': torch.cat([text_gauss_m, t['vis_m']], -1),
                     'textaudiovisual': torch.cat([text_gauss_m, t['aud_m'], t['vis_m']], -1)}
        We_t = t['We']
        a = 1e-3

        def word_fn(latents, word_weights, sent_embeddings, mask):  # simplesif.py:527-537
            return ref_losses.get_word_log_prob_angular2(latents, We_t, word_weights,
                                                         sent_embeddings, mask, a)
        args = dict(cfg['args'])
        total = ref_losses.get_log_prob_matrix(args, lat, out_heads, data, masks, word_fn)
        loss = (-total).mean()                                      # simplesif.py:129-134
        loss.backward()
        res = dict(total=total.detach().numpy(), loss=loss.detach().numpy(),
                   grad_latents=lat.grad.numpy())
        for mod, dd in out_heads.items():
            res['mu_' + mod] = dd['mu'].detach().numpy()
            res['sigma_' + mod] = dd['sigma'].detach().numpy()
            res['lp_' + mod] = ref_losses.get_normal_log_prob(
                dd['mu'].unsqueeze(1), dd['sigma'].unsqueeze(1), data[mod], masks[mod]).detach().numpy()
            for nm in ('mu', 'log_sigma'):
                gW = model.embed2out[mod][nm].weight.grad.numpy()
                # large case: keep 8 rows of each weight gradient + a checksum of all of it
                res['gW_%s_%s' % (nm, mod)] = gW if gW.size <= 4096 else gW[:8]
                res['gWsum_%s_%s' % (nm, mod)] = cases.checksum(gW)
                res['gb_%s_%s' % (nm, mod)] = model.embed2out[mod][nm].bias.grad.numpy()
        if model.norm is not None:
            res['g_norm_w'] = model.norm.weight.grad.numpy()
            res['g_norm_b'] = model.norm.bias.grad.numpy()
        res['word_lp'] = word_fn(lat, data['text_weights'], data['text'], masks['text']).detach().numpy()
        lat2 = torch.tensor(c['latents'], requires_grad=True)
        word_fn(lat2, data['text_weights'], data['text'], masks['text']).sum().backward()
        res['word_grad'] = lat2.grad.numpy()
        res['inputs_sum'] = cases.checksum(np.concatenate(
            [c[k].ravel() for k in sorted(c) if isinstance(c[k], np.ndarray)]))
        save('mmb_%s.npz' % tag, **res)


def golden_optimize_latents():
    """Run the reference's own optimize_latents (simplesif.py:49-162, executed from the source
    file -- the module itself cannot be imported) on a small MOSI-shaped
    problem with a fixed batch order, for train=True (SGD on latents + heads) and train=False."""
    import time
    import torch.optim as optim
    from torch.utils.data import DataLoader
    import utils_stub
    src = open(os.path.join(REF, 'simplesif.py')).read().split('\n')
    code = '\n'.join(src[48:162])
    ns = {'torch': torch, 'optim': optim, 'time': time, 'np': np,
          'get_log_prob_matrix': ref_losses.get_log_prob_matrix}
    exec(compile(code, 'reference simplesif.py:49-162', 'exec'), ns)
    ref_optimize_latents = ns['optimize_latents']
    out = {}
    for tag, cfg in cases.OPT_CASES.items():
        c = cases.mmb_inputs(**cfg['inputs'])
        torch.manual_seed(0)
        model = ref_models.AudioVisualGeneratorMultimodal(c['d'], c['A'], c['Vd'], norm=cfg['inputs']['norm'],
                                                          frozen_weights=False, unimodal=cfg['inputs']['unimodal'])
        cases.load_heads(model, c['heads'], c.get('norm_params'))
        ds = utils_stub.MMData(c['text'], c['aud'], c['vis'], {'text': c['text_m'], 'covarep': c['aud_m'],
                                                                'facet': c['vis_m']}, c['text_w'], torch.device('cpu'))
        loader = DataLoader(ds, batch_size=cfg['batch'], shuffle=False)
        We_t = torch.tensor(c['We'])

        def word_fn(latents, word_weights, sent_embeddings, mask):
            return ref_losses.get_word_log_prob_angular2(latents, We_t, word_weights, sent_embeddings, mask, 1e-3)
        args = dict(cfg['args'])
        emb, (losses, _) = ref_optimize_latents(args, cfg['train'], model, c['latents'], loader, cfg['epochs'],
                                                cfg['lr'], word_fn, torch.device('cpu'), verbose=False)
        out[tag + '_emb'] = emb.detach().numpy()
        out[tag + '_losses'] = np.array(losses)
        out[tag + '_Wmu_audio'] = model.embed2out['audio']['mu'].weight.detach().numpy()
    save('optimize_latents.npz', **out)


def golden_closed_form():
    """reference sif2.estimate_embedding_overall_gpu2 (the closed-form latents timed under
    --time_test, simplesif.py:808-880) on seeded splits; sif2 imports utils -> h5py is stubbed."""
    import utils_stub  # noqa: F401  (installs the h5py stub)
    import sif2 as ref_sif2
    assert ref_sif2.__file__.startswith(REF)
    out = {}
    for tag, cfg in cases.CLOSED_FORM_CASES.items():
        c = cases.mmb_inputs(**cfg)
        model = ref_models.AudioVisualGeneratorMultimodal(c['d'], c['A'], c['Vd'], norm=None, frozen_weights=True,
                                                          unimodal=False)
        cases.load_heads(model, c['heads'])
        data = {k: torch.tensor(v) for k, v in cases.closed_form_data(c).items()}
        networks = {k: (model.embed2out[k]['mu'], model.embed2out[k]['log_sigma']) for k in cases.CLOSED_FORM_KEYS}
        with torch.no_grad():
            cs = ref_sif2.estimate_embedding_overall_gpu2(data, {k: None for k in data}, networks, torch.tensor(c['text_w']),
                                                         torch.tensor(c['text']))
        out[tag + '_cs'] = cs.numpy()
        out[tag + '_check'] = cases.checksum(c['text']) + cases.checksum(c['aud'])
    save('closed_form.npz', **out)


def golden_utils():
    """Reference utils.py preprocessing (normalize_data 155-191, add_positional_embeddings
    130-153, quirks included) on a small seeded split."""
    import utils_stub
    split = cases.raw_split()
    pos = utils_stub.add_positional_embeddings({'pos_embed_dim': 4}, split['covarep'].copy())
    norm, masks = utils_stub.normalize_data({k: v.copy() for k, v in split.items()})
    save('utils.npz', pos=pos, covarep=norm['covarep'], facet=norm['facet'], m_covarep=masks['covarep'],
         m_facet=masks['facet'])


def golden_downstream():
    """The script path end to end through the UNMODIFIED reference on small labelled splits:
    sif.get_sentence_embeddings per split -> id expansion / positional columns (simplesif.py:296-399,
    restated from main(), which cannot be imported) -> MMData / DataLoader (442-459) -> the e2e block
    (simplesif.py:671-806, executed from the source file) or the two-stage branch (optimize_latents,
    588-610) -> sentiment_model.train_sentiment_for_latents -> test_results_{before,after}.json."""
    import json
    import tempfile
    import textwrap
    import time
    import torch.nn as nn
    import torch.optim as optim
    from torch.utils.data import DataLoader
    import utils_stub
    import sentiment_model as ref_sm
    assert ref_sm.__file__.startswith(REF)
    src = open(os.path.join(REF, 'simplesif.py')).read().split('\n')
    ns = {'torch': torch, 'optim': optim, 'time': time, 'np': np, 'nn': nn,
          'get_log_prob_matrix': ref_losses.get_log_prob_matrix}
    exec(compile('\n'.join(src[35:47]), 'reference simplesif.py:36-47', 'exec'), ns)
    exec(compile('\n'.join(src[48:162]), 'reference simplesif.py:49-162', 'exec'), ns)
    e2e_block = textwrap.dedent('\n'.join(src[670:806]))
    out = {}
    device = torch.device('cpu')
    for tag, cfg in cases.DOWNSTREAM_CASES.items():
        args = dict(cfg['args'])
        We, weights, splits, masks = cases.downstream_inputs(**cfg)
        torch.manual_seed(cfg['seed'])
        id_key = 'text' if args['dataset'] == 'mosi' else 'text_id'
        for s, m in zip(splits, masks):
            ns['update_masks'](m, s[id_key], We.shape[-1])
        emb = [ref_sif.get_sentence_embeddings(We, weights, s[id_key]) for s in splits]
        w_t = torch.tensor(weights, device=device, dtype=torch.float32)
        we_t = torch.tensor(We, device=device, dtype=torch.float32)
        for s, m in zip(splits, masks):                       # simplesif.py:319-399
            if args['dataset'] == 'mosi':
                s['text_id'] = s['text']
            else:
                s['text_align'] = s['text']
                ns['update_masks_vect'](m, s['text_align'], 'text_align')
            s['text'] = we_t[s['text_id']]
            s['text_weights'] = w_t[s['text_id']]
            n_points, seq_len = m['covarep'].shape[:2]
            ext = np.ones((n_points, seq_len, args['pos_embed_dim']), dtype=np.int64)
            for k in ('covarep', 'facet'):
                s[k] = utils_stub.add_positional_embeddings(args, s[k])
                m[k] = np.concatenate([m[k], ext], axis=-1)
        train, valid, test = splits

        def dataset(s, m):
            if args['dataset'] == 'mosi':
                return utils_stub.MMData(s['text'], s['covarep'], s['facet'], m, s['text_weights'], device)
            return utils_stub.MMDataExtra(s['text'], s['covarep'], s['facet'], m, s['text_weights'],
                                          s['text_align'], device)
        bs = args['batch_size']
        dataloader = DataLoader(dataset(train, masks[0]), batch_size=bs, shuffle=True)
        valid_dataloader = DataLoader(dataset(valid, masks[1]), batch_size=bs * 8)
        test_dataloader = DataLoader(dataset(test, masks[2]), batch_size=bs * 8)

        def get_word_log_prob2(latents, word_weights, sent_embeddings, mask):   # simplesif.py:527-537
            return ref_losses.get_word_log_prob_angular2(latents, we_t, word_weights, sent_embeddings, mask, 1e-3)
        sentiment_data = (train['label'], valid['label'], test['label'])
        d, A, Vd = train['text'].shape[-1], train['covarep'].shape[-1], train['facet'].shape[-1]
        if args['e2e']:
            env = dict(ns)
            env.update(AudioVisualGeneratorMultimodal=ref_models.AudioVisualGeneratorMultimodal, EMBEDDING_DIM=d,
                       AUDIO_DIM=A, VISUAL_DIM=Vd, args=args, device=device, train=train,
                       SentimentModel=ref_sm.SentimentModel, train_embedding=emb[0], valid_embedding=emb[1],
                       test_embedding=emb[2], dataloader=dataloader, valid_dataloader=valid_dataloader,
                       test_dataloader=test_dataloader, senti_train_data=ref_sm.SentimentData(train['label'], device),
                       get_word_log_prob2=get_word_log_prob2, sentiment_train_idxes=None,
                       senti_mask=torch.zeros(len(train['label']), device=device))
            exec(compile(e2e_block, 'reference simplesif.py:671-806', 'exec'), env)
            train_embed, valid_embed, test_embed = env['train_embed'], env['valid_embed'], env['test_embed']
            train_losses = env['train_losses']
        else:
            gen_model = ref_models.AudioVisualGeneratorMultimodal(d, A, Vd, norm=args['norm'],
                                                                  frozen_weights=args['freeze_weights'],
                                                                  unimodal=args['unimodal']).to(device)
            opt = ns['optimize_latents']
            train_embed, (train_losses, _) = opt(args, True, gen_model, emb[0], dataloader, args['n_epochs'],
                                                 args['lr'], get_word_log_prob2, device,
                                                 validation_data=(emb[1], valid_dataloader))
            valid_embed, _ = opt(args, False, gen_model, emb[1], valid_dataloader, args['n_epochs'], args['lr'],
                                 get_word_log_prob2, device)
            test_embed, _ = opt(args, False, gen_model, emb[2], test_dataloader, args['n_epochs'], args['lr'],
                                get_word_log_prob2, device)
        with tempfile.TemporaryDirectory() as tmp:
            ref_sm.train_sentiment_for_latents(args, (train_embed.detach(), valid_embed, test_embed), sentiment_data,
                                               device, train_idxes=None, model_save_path=tmp)
            before = json.load(open(os.path.join(tmp, 'test_results_before.json')))
            after = json.load(open(os.path.join(tmp, 'test_results_after.json')))
        for nm, r in (('before', before), ('after', after)):
            for k in ('mae', 'corr', 'mult_acc', 'f_score', 'accuracy'):
                if k in r:
                    out['%s_%s_%s' % (tag, nm, k)] = np.asarray(r[k], dtype=np.float64)
        out[tag + '_sif_test'] = emb[2][:8]
        out[tag + '_train_embed'] = train_embed.detach().numpy()[:8]
        out[tag + '_test_embed'] = test_embed.detach().numpy()[:8]
        out[tag + '_train_embed_sum'] = cases.checksum(train_embed.detach().numpy())
        out[tag + '_test_embed_sum'] = cases.checksum(test_embed.detach().numpy())
        out[tag + '_train_losses'] = np.asarray(train_losses, dtype=np.float64)
        out[tag + '_inputs_sum'] = cases.checksum(np.concatenate([We.ravel()] + [s['covarep'].ravel() for s in splits]))
    save('downstream.npz', **out)


def golden_sentiment():
    """reference sentiment_model.train_sentiment_for_latents on fixed latents (CPU)."""
    import json
    import tempfile
    import sentiment_model as ref_sm
    assert ref_sm.__file__.startswith(REF)
    out = {}
    for tag, cfg in cases.SENTIMENT_CASES.items():
        args, lat, labs = cases.sentiment_inputs(**cfg)
        torch.manual_seed(cfg['seed'])
        with tempfile.TemporaryDirectory() as tmp:
            ref_sm.train_sentiment_for_latents(args, tuple(torch.tensor(x) for x in lat), tuple(labs),
                                               torch.device('cpu'), model_save_path=tmp)
            for nm in ('before', 'after'):
                r = json.load(open(os.path.join(tmp, 'test_results_%s.json' % nm)))
                for k in ('mae', 'corr', 'mult_acc', 'f_score', 'accuracy'):
                    if k in r:
                        out['%s_%s_%s' % (tag, nm, k)] = np.asarray(r[k], dtype=np.float64)
            out[tag + '_valid_losses'] = np.array([float(x) for x in open(os.path.join(tmp, 'senti_valid_loss.txt'))])
        out[tag + '_next_draw'] = torch.rand(1).numpy()      # the global generator ends in the same state
    save('sentiment.npz', **out)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'sentiment':
    golden_sentiment()
    sys.exit(0)


def golden_pom_full():
    """BASELINE config 3 at full size on the real fixtures: both POM splits in the tree (valid 100 x 1089,
    test 203 x 1357, ids right-padded with 0 whose weight is 1.0) and the real pom_word_weights through
    the reference's sif.get_sentence_embeddings; only the 300-d table is synthetic."""
    pom_w = np.load(os.path.join(REF, 'pom', 'pom_word_weights.npy')).squeeze()
    We = cases.table(7763, 300, seed=11)
    out = dict(weights=pom_w, table_sum=cases.checksum(We))
    for split in ('valid', 'test'):
        ids = np.load(os.path.join(REF, 'pom', 'pom_%s_ids.npy' % split))
        assert ids.max() < 2 ** 15 and ids.min() >= 0
        w = ref_sf.seq2weight(ids, np.ones(ids.shape), pom_w)
        avg = ref_sf.get_weighted_average(We, ids, w)
        emb = ref_sif.get_sentence_embeddings(We, pom_w, ids)
        out.update({split + '_ids': ids.astype(np.int16), split + '_w_nonzero': np.count_nonzero(w, axis=1),
                    split + '_avg': avg.astype(np.float32), split + '_emb': emb.astype(np.float32),
                    split + '_emb_sum': cases.checksum(emb), split + '_pc': ref_sf.compute_pc(avg, 1)})
    save('sif_pom_full.npz', **out)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'pom_full':
    golden_pom_full()
    sys.exit(0)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'optimize_latents':
    golden_optimize_latents()
    sys.exit(0)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'downstream':
    golden_downstream()
    sys.exit(0)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'closed_form':
    golden_closed_form()
    sys.exit(0)

if __name__ == '__main__':
    main()
    golden_optimize_latents()
    golden_utils()
    golden_closed_form()
    golden_downstream()
    golden_sentiment()
    golden_pom_full()
