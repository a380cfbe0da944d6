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
from .database import Database, read_file_list, write_flat_ip_index_from_file
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


# Files are processed in bounded batches (the reference streams file by file, builder.py:75-103): a 10 M-row
# database is ~80 GB of int16 PCM, which must never sit in host memory at once.
BATCH_PCM_BYTES = 1 << 30      # ~1 GB of decoded PCM per GPU call
BATCH_FILES = 4096


def iter_extract(ex, files, frame_shift_mul, max_bytes=None, max_files=None):
    """Yields (first file index, emb [n, d] fp32, counts [n_files_in_batch]) for consecutive batches of `files`.
    Fingerprints are in list order; a count of 0 = unreadable file (builder.py:82-86).  Mono 16-bit files at the
    model rate take the fused PCM path, everything else the GPU ingest (resample + mono mix); one batched GPU call
    per kind and batch."""
    sr = ex.params['sample_rate']
    max_bytes = BATCH_PCM_BYTES if max_bytes is None else max_bytes
    max_files = BATCH_FILES if max_files is None else max_files
    i = 0
    while i < len(files):
        i0, kinds, datas, nbytes = i, [], [], 0
        while i < len(files) and len(kinds) < max_files and (nbytes < max_bytes or not kinds):
            try:
                kind, data = musicdata.read_wav_pcm16(files[i], sr)
                nbytes += data.nbytes if kind == 'pcm16' else data[0].nbytes
            except Exception as e:  # noqa: BLE001  (musicdata.py:95-101: log and yield zero segments)
                print('load %s error! %s' % (files[i], e))
                kind, data = 'error', None
            kinds.append(kind)
            datas.append(data)
            i += 1
        counts = np.zeros(len(kinds), np.int64)
        parts = [None] * len(kinds)
        pcm_ids = [j for j, k in enumerate(kinds) if k == 'pcm16']
        if pcm_ids:
            off = np.concatenate([[0], np.cumsum([len(datas[j]) for j in pcm_ids])]).astype(np.int64)
            pcm = np.concatenate([datas[j] for j in pcm_ids]) if off[-1] else np.zeros(0, np.int16)
            z, cnt = ex.extract_pcm16(pcm, off, frame_shift_mul=frame_shift_mul)
            pos = np.concatenate([[0], np.cumsum(cnt)])
            for jj, j in enumerate(pcm_ids):
                parts[j] = z[pos[jj]:pos[jj + 1]]
                counts[j] = cnt[jj]
        wav_ids = [j for j, k in enumerate(kinds) if k == 'wav']      # multi-channel / other rates: GPU ingest
        if wav_ids:
            z, cnt = ex.extract_wavs([datas[j] for j in wav_ids], frame_shift_mul=frame_shift_mul)
            pos = np.concatenate([[0], np.cumsum(cnt)])
            for jj, j in enumerate(wav_ids):
                parts[j] = z[pos[jj]:pos[jj + 1]]
                counts[j] = cnt[jj]
        emb = [q for q in parts if q is not None and len(q)]
        yield i0, (np.concatenate(emb) if emb else np.zeros((0, ex.d), np.float32)), counts


def extract_files(ex, files, frame_shift_mul, log=None):
    """All fingerprints at once (small lists / tests); the command lines stream through iter_extract."""
    embs, counts = [], []
    for _, e, c in iter_extract(ex, files, frame_shift_mul):
        embs.append(e)
        counts.append(c)
    emb = np.concatenate(embs) if embs else np.zeros((0, ex.d), np.float32)
    return emb, (np.concatenate(counts) if counts else np.zeros(0, np.int64))


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
    factory = params.get('indexer', {}).get('index_factory', 'Flat')
    if factory != 'Flat':
        # builder.py:111-130 trains a faiss index of this type; this build searches exactly (BASELINE configs 3/4)
        print('warning: index_factory %r is not built here; landmarkValue is written as an exact inner-product '
              'index (IndexFlatIP) and searched by brute force' % factory)
    t0 = time.time()
    total, counts = 0, []
    with open(os.path.join(dir_for_db, 'embeddings'), 'wb') as femb:            # builder.py:71,99: appended per batch
        for _, emb, cnt in iter_extract(ex, files, 1):                          # builder.py:64: always fsm = 1
            femb.write(np.ascontiguousarray(emb, dtype=np.float32).tobytes())
            total += emb.shape[0]
            counts.append(cnt)
    counts = np.concatenate(counts) if counts else np.zeros(0, np.int64)
    print('total', total, 'embeddings (%.3fs)' % (time.time() - t0))
    if total == 0:
        print('The database is empty!')
    print('writing database')
    write_flat_ip_index_from_file(os.path.join(dir_for_db, 'landmarkValue'), os.path.join(dir_for_db, 'embeddings'),
                                  total, ex.d)                                  # builder.py:135-136
    counts.astype(np.int32).tofile(os.path.join(dir_for_db, 'landmarkKey'))         # builder.py:138-139
    shutil.copyfile(file_list_for_db, os.path.join(dir_for_db, 'songList.txt'))     # builder.py:141
    shutil.copyfile(cfg_path, os.path.join(dir_for_db, 'configs.json'))             # builder.py:144
    shutil.copyfile(os.path.join(params['model_dir'], 'model.pt'), os.path.join(dir_for_db, 'model.pt'))
    return 0


class _ResultWriter:
    """matcher.py:40-42,80-84,157-166: the three result files, written row by row as batches finish (the .bin
    grows by n_songs * 8 bytes per query and is never held in memory)."""

    def __init__(self, result_file, db):
        self.db = db
        self.fout = open(result_file, 'w', encoding='utf8', newline='\n')
        self.fout2 = open(os.path.splitext(result_file)[0] + '_detail.csv', 'w', encoding='utf8', newline='\n')
        self.fbin = open(result_file + '.bin', 'wb')
        self.w = csv.writer(self.fout2)
        self.w.writerow(['query', 'answer', 'score', 'time', 'part_scores'])         # matcher.py:84
        self.n_songs = len(db.songList)

    def error(self, name):                                                           # matcher.py:94-107
        self.fout.write('%s\t%s\n' % (name, 'error'))
        self.w.writerow([name, 'error', -1e999, 0])
        self.fbin.write(np.zeros([self.n_songs, 2], dtype=np.float32).tobytes())

    def row(self, name, score, song, tim, song_scores):
        ans = self.db.songList[song]                                                 # matcher.py:138 (-1 -> last)
        self.fout.write('%s\t%s\n' % (name, ans))
        self.w.writerow([name, ans, score, tim])
        self.fbin.write(np.ascontiguousarray(song_scores, dtype=np.float32).tobytes())

    def close(self):
        for f in (self.fout, self.fout2, self.fbin):
            f.close()


def _match(db, names, emb, counts, out, batch=256):
    """Search + sequence score for the query files `names` (fingerprints `emb`, `counts[i]` rows each); rows go
    to the writer `out` in list order."""
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    i = 0
    while i < len(names):
        if counts[i] <= 0:
            out.error(names[i])
            i += 1
            continue
        ids = []
        while i < len(names) and counts[i] > 0 and len(ids) < batch:
            ids.append(i)
            i += 1
        q = np.concatenate([emb[pos[j]:pos[j + 1]] for j in ids])
        lens = np.array([counts[j] for j in ids], np.int64)
        qi = np.stack([np.concatenate([[0], np.cumsum(lens)[:-1]]), lens], axis=1)
        s, g, t, ss = db.query_batch(q, qi, want_song_scores=True)
        for jj, j in enumerate(ids):
            out.row(names[j], float(s[jj]), int(g[jj]), float(t[jj]), ss[jj])


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
    out = _ResultWriter(result_file, db)
    try:
        for i0, emb, counts in iter_extract(ex, names, fsm):
            _match(db, names[i0:i0 + len(counts)], emb, counts, out)
    finally:
        out.close()
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
    os.makedirs(out_dir, exist_ok=True)
    total, counts = 0, []
    with open(os.path.join(out_dir, 'query_embeddings'), 'wb') as femb:             # extractemb.py:83
        for _, emb, cnt in iter_extract(ex, names, fsm):
            femb.write(np.ascontiguousarray(emb, dtype=np.float32).tobytes())
            total += emb.shape[0]
            counts.append(cnt)
    counts = np.concatenate(counts) if counts else np.zeros(0, np.int64)
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    np.stack([pos[:-1], counts.astype(np.int64)], axis=1).tofile(os.path.join(out_dir, 'query_index'))  # :85
    print('total', total, 'embeddings')
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
    emb = np.memmap(os.path.join(dir_for_query, 'query_embeddings'), dtype=np.float32, mode='r').reshape([-1, d]) \
        if os.path.getsize(os.path.join(dir_for_query, 'query_embeddings')) else np.zeros((0, d), np.float32)
    qidx = np.fromfile(os.path.join(dir_for_query, 'query_index'), dtype=np.int64).reshape([-1, 2])
    # matchemb.py:63-65 slices [start, start+len) per file; files that failed to load have len 0.  The embedding
    # file is memory-mapped and walked in batches of query files.
    out = _ResultWriter(result_file, db)
    try:
        step = 4096
        for i0 in range(0, len(names), step):
            part = qidx[i0:i0 + step]
            flat = np.concatenate([emb[s:s + l] for s, l in part]) if len(part) else emb[:0]
            _match(db, names[i0:i0 + step], flat, part[:, 1], out)
    finally:
        out.close()
    return 0
