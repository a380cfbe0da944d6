"""Drop-in command lines: same argv, same files on disk as the reference's builder.py / matcher.py /
extractemb.py / matchemb.py, with stages 1-3 on the GPU.

    python builder.py    <music list> <db dir> [config.json | model dir]        builder.py:30-44
    python matcher.py    <query list> <db dir> <result file>                    matcher.py:34-44
    python extractemb.py <query list> <db dir> <output embedding dir>           extractemb.py:20-27
    python matchemb.py   <query embedding dir> <db dir> <result file>           matchemb.py:19-31

Database directory (builder.py:70-71,136-148): embeddings (raw fp32 [N,d]), landmarkValue (here always a
faiss-format IndexFlatIP file: the BASELINE search is brute force), landmarkKey (int32 segments per song),
songList.txt, configs.json, model.pt.  Results (matcher.py:40-42,157-166): <result> TSV "query\\tanswer",
<stem>_detail.csv [query, answer, score, time], <result>.bin fp32 [n_queries, n_songs, 2].
"""
import csv
import os
import shutil
import sys
import time

import numpy as np
import torch

from . import synth
from .database import Database, read_file_list, write_flat_ip_index
from .datautil import musicdata
from .extract import Extractor


def _read_params(configs):
    """builder.py:35-44: a JSON file, or a model directory holding configs.json + model.pt."""
    if os.path.isdir(configs):
        path = os.path.join(configs, 'configs.json')
        params = synth.read_config(path)
        params['model_dir'] = configs
        return params, path
    return synth.read_config(configs), configs


def _load_state(model_dir):
    sd = torch.load(os.path.join(model_dir, 'model.pt'), map_location='cpu')
    return {k: v.float() for k, v in sd.items()}


def extract_files(ex, files, frame_shift_mul, log=None):
    """Fingerprints of every file, concatenated in list order, + segments per file (0 = unreadable,
    builder.py:82-86).  Mono 16-bit files at the model rate take the fused PCM path in ONE batched call."""
    sr = ex.params['sample_rate']
    kinds, datas = [], []
    for f in files:
        try:
            kind, data = musicdata.read_wav_pcm16(f, sr)
        except Exception as e:  # noqa: BLE001  (musicdata.py:95-101: log and yield zero segments)
            print('load %s error! %s' % (f, e))
            kind, data = 'error', None
        kinds.append(kind)
        datas.append(data)
    counts = np.zeros(len(files), np.int64)
    parts = [None] * len(files)
    pcm_ids = [i for i, k in enumerate(kinds) if k == 'pcm16']
    if pcm_ids:
        off = np.concatenate([[0], np.cumsum([len(datas[i]) for i in pcm_ids])]).astype(np.int64)
        pcm = np.concatenate([datas[i] for i in pcm_ids]) if off[-1] else np.zeros(0, np.int16)
        z, cnt = ex.extract_pcm16(pcm, off, frame_shift_mul=frame_shift_mul)
        pos = np.concatenate([[0], np.cumsum(cnt)])
        for j, i in enumerate(pcm_ids):
            parts[i] = z[pos[j]:pos[j + 1]]
            counts[i] = cnt[j]
    for i, k in enumerate(kinds):
        if k == 'float':
            rows = musicdata.frame_float(datas[i], ex.seg_len, ex.hop // frame_shift_mul)
            parts[i] = ex.extract_segments(rows)
            counts[i] = rows.shape[0]
    emb = [p for p in parts if p is not None and len(p)]
    emb = np.concatenate(emb) if emb else np.zeros((0, ex.d), np.float32)
    return emb, counts


def builder_main(argv):
    if len(argv) < 3:
        print('Usage: python %s <music list file> <db location>' % argv[0])
        return 0
    file_list_for_db, dir_for_db = argv[1], argv[2]
    params, cfg_path = _read_params(argv[3] if len(argv) >= 4 else os.path.join(synth.REPO, 'configs', 'default.json'))
    print('loading model...')
    ex = Extractor(params, _load_state(params['model_dir']))
    print('model loaded')
    files = read_file_list(file_list_for_db)
    os.makedirs(dir_for_db, exist_ok=True)
    t0 = time.time()
    emb, counts = extract_files(ex, files, 1)          # builder.py:64: the database is always built at fsm = 1
    print('total', emb.shape[0], 'embeddings (%.3fs)' % (time.time() - t0))
    if emb.shape[0] == 0:
        print('The database is empty!')
    emb.astype(np.float32).tofile(os.path.join(dir_for_db, 'embeddings'))           # builder.py:99
    print('writing database')
    write_flat_ip_index(os.path.join(dir_for_db, 'landmarkValue'), emb)             # builder.py:135-136
    counts.astype(np.int32).tofile(os.path.join(dir_for_db, 'landmarkKey'))         # builder.py:138-139
    shutil.copyfile(file_list_for_db, os.path.join(dir_for_db, 'songList.txt'))     # builder.py:141
    shutil.copyfile(cfg_path, os.path.join(dir_for_db, 'configs.json'))             # builder.py:144
    shutil.copyfile(os.path.join(params['model_dir'], 'model.pt'), os.path.join(dir_for_db, 'model.pt'))
    return 0


def _write_results(result_file, names, db, scores, songs, times, song_scores):
    result_file2 = os.path.splitext(result_file)[0] + '_detail.csv'
    with open(result_file, 'w', encoding='utf8', newline='\n') as fout, \
            open(result_file2, 'w', encoding='utf8', newline='\n') as fout2, \
            open(result_file + '.bin', 'wb') as fbin:
        w = csv.writer(fout2)
        w.writerow(['query', 'answer', 'score', 'time', 'part_scores'])              # matcher.py:84
        n_songs = len(db.songList)
        for i, name in enumerate(names):
            if songs[i] is None:                                                     # matcher.py:94-107
                fout.write('%s\t%s\n' % (name, 'error'))
                w.writerow([name, 'error', -1e999, 0])
                fbin.write(np.zeros([n_songs, 2], dtype=np.float32).tobytes())
                continue
            ans = db.songList[songs[i]]                                              # matcher.py:138 (-1 -> last)
            fout.write('%s\t%s\n' % (name, ans))
            w.writerow([name, ans, scores[i], times[i]])
            fbin.write(song_scores[i].astype(np.float32).tobytes())


def _match(db, names, emb, counts, result_file, batch=256):
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    scores, songs, times, sscores = [None] * len(names), [None] * len(names), [None] * len(names), [None] * len(names)
    ok = [i for i in range(len(names)) if counts[i] > 0]
    for b0 in range(0, len(ok), batch):
        ids = ok[b0:b0 + batch]
        q = np.concatenate([emb[pos[i]:pos[i + 1]] for i in ids])
        lens = np.array([counts[i] for i in ids], np.int64)
        qi = np.stack([np.concatenate([[0], np.cumsum(lens)[:-1]]), lens], axis=1)
        s, g, t, ss = db.query_batch(q, qi, want_song_scores=True)
        for j, i in enumerate(ids):
            scores[i], songs[i], times[i], sscores[i] = float(s[j]), int(g[j]), float(t[j]), ss[j]
    _write_results(result_file, names, db, scores, songs, times, sscores)


def matcher_main(argv):
    if len(argv) < 4:
        print('Usage: python %s <query list> <database dir> <result file>' % argv[0])
        return 0
    file_list_for_query, dir_for_db, result_file = argv[1], argv[2], argv[3]
    params = synth.read_config(os.path.join(dir_for_db, 'configs.json'))            # matcher.py:43-44
    fsm = params['indexer'].get('frame_shift_mul', 1)
    print('loading model...')
    ex = Extractor(params, _load_state(dir_for_db))                                 # matcher.py:60-61
    print('model loaded')
    print('loading database...')
    db = Database(dir_for_db, params['indexer'], params['hop_size'])                # matcher.py:65
    print('database loaded')
    names = read_file_list(file_list_for_query)
    t0 = time.time()
    emb, counts = extract_files(ex, names, fsm)
    _match(db, names, emb, counts, result_file)
    print('total query time %.6fs' % (time.time() - t0))
    return 0


def extractemb_main(argv):
    if len(argv) < 4:
        print('Usage: python %s <query list> <database dir> <output embedding dir>' % argv[0])
        return 0
    file_list_for_query, dir_for_db, out_dir = argv[1], argv[2], argv[3]
    cfg = os.path.join(dir_for_db, 'configs.json')
    params = synth.read_config(cfg)
    fsm = params['indexer'].get('frame_shift_mul', 1)
    ex = Extractor(params, _load_state(dir_for_db))
    names = read_file_list(file_list_for_query)
    emb, counts = extract_files(ex, names, fsm)
    os.makedirs(out_dir, exist_ok=True)
    emb.astype(np.float32).tofile(os.path.join(out_dir, 'query_embeddings'))        # extractemb.py:83
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    np.stack([pos[:-1], counts.astype(np.int64)], axis=1).tofile(os.path.join(out_dir, 'query_index'))  # :85
    print('total', emb.shape[0], 'embeddings')
    shutil.copyfile(file_list_for_query, os.path.join(out_dir, 'queryList.txt'))    # extractemb.py:90
    shutil.copyfile(cfg, os.path.join(out_dir, 'configs.json'))
    return 0


def matchemb_main(argv):
    if len(argv) < 4:
        print('Usage: python %s <query embedding dir> <database dir> <result file>' % argv[0])
        return 0
    dir_for_query, dir_for_db, result_file = argv[1], argv[2], argv[3]
    params = synth.read_config(os.path.join(dir_for_db, 'configs.json'))
    names = read_file_list(os.path.join(dir_for_query, 'queryList.txt'))
    d = params['model']['d']
    db = Database(dir_for_db, params['indexer'], params['hop_size'])
    emb = np.fromfile(os.path.join(dir_for_query, 'query_embeddings'), dtype=np.float32).reshape([-1, d])
    qidx = np.fromfile(os.path.join(dir_for_query, 'query_index'), dtype=np.int64).reshape([-1, 2])
    # matchemb.py:63-65 slices [start, start+len) per file; files that failed to load have len 0
    flat = np.concatenate([emb[s:s + l] for s, l in qidx]) if len(qidx) else emb[:0]
    _match(db, names, flat, qidx[:, 1], result_file)
    return 0
