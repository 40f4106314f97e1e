"""Continuous-batching scheduler for streaming recognition.

The reference serves one stream per predictor (`PPASRPredictor.predict_stream`, predict.py:232-337; the web-socket loop of
infer_server.py:103-156 owns a single global predictor). Here many sessions share one engine: every session owns a cache
slot on the device (attention K/V + conv caches, its own position), buffers its incoming fbank frames with the reference's
window rule (67-frame windows, stride 64, 3 frames of context kept; predict.py:277-300), and `step()` runs ONE batched
encoder chunk for all sessions that have a full window, whatever their positions (`ppasr_b200_sessions_step`). Decoding is
the reference's chunk-wise greedy decoder per session (decoders/ctc_greedy_decoder.py:52-89). Sessions take fbank frames
(`feed`) or raw audio (`feed_audio`: the per-session copy of predict_stream's audio bookkeeping, predict.py:250-274)."""
import numpy as np

from ..decoders.ctc_greedy_decoder import greedy_decoder_chunk

DECODING_CHUNK = 16
CONTEXT = 7
SUBSAMPLING = 4
WINDOW = (DECODING_CHUNK - 1) * SUBSAMPLING + CONTEXT  # 67
STRIDE = SUBSAMPLING * DECODING_CHUNK                  # 64


class _Session:
    __slots__ = ("slot", "feat", "ended", "ids", "probs", "text", "score", "wav")

    def __init__(self, slot, n_mels):
        self.slot = slot
        self.wav = None     # audio not yet turned into frames (feed_audio), dB-normalised in place like predict.py:262-274
        self.feat = np.zeros((0, n_mels), dtype=np.float32)
        self.ended = False
        self.ids = None
        self.probs = None
        self.text = ""
        self.score = 0.0


class StreamScheduler:
    def __init__(self, predictor, vocabulary, max_sessions=64, required_cache_size=-1):
        """predictor: ppasr_b200 InferencePredictor of a streaming conformer."""
        if predictor.use_model != "conformer" or not predictor.streaming:
            raise Exception("StreamScheduler needs a streaming conformer predictor")
        self.engine = predictor.engine
        self.vocab = vocabulary
        self.n_mels = predictor.model_config.input_dim
        self.required = int(required_cache_size)
        self.max_sessions = int(max_sessions)
        self.engine.sessions_init(self.max_sessions)
        self._free = list(range(self.max_sessions - 1, -1, -1))
        conf = getattr(predictor, "configs", None)
        self._preprocess_conf = (conf.get("preprocess_conf") if isinstance(conf, dict) else getattr(conf, "preprocess_conf", None))
        self._featurizer = None
        self._sessions = {}
        self._next_sid = 0

    def open(self):
        if not self._free:
            raise Exception("no free stream slot")
        slot = self._free.pop()
        self.engine.sessions_reset(slot)
        sid = self._next_sid
        self._next_sid += 1
        self._sessions[sid] = _Session(slot, self.n_mels)
        return sid

    def feed(self, sid, feats, is_end=False):
        """feats: [n, n_mels] new fbank frames of the session."""
        s = self._sessions[sid]
        s.feat = np.concatenate([s.feat, np.asarray(feats, dtype=np.float32).reshape(-1, self.n_mels)], axis=0)
        s.ended = s.ended or bool(is_end)

    def feed_audio(self, sid, audio_data, is_end=False, channels=1, samp_width=2, sample_rate=16000):
        """Raw audio of one session, exactly the front half of PPASRPredictor.predict_stream (predict.py:250-274): ndarray
        (float or integer) or headerless PCM bytes are appended to the session's remaining audio, ALL of it is featurised (the
        per-call dB normalisation is applied in place, so the unconsumed tail stays scaled), the consumed samples are dropped
        and the new frames are queued with feed()."""
        s = self._sessions[sid]
        fz = self._audio_featurizer()
        if isinstance(audio_data, np.ndarray):
            x = fz.to_float32(audio_data)
        elif isinstance(audio_data, bytes):
            x = fz.from_pcm_bytes(audio_data, channels=channels, samp_width=samp_width)
        else:
            raise Exception(f'不支持该数据类型，当前数据类型为：{type(audio_data)}')
        s.wav = x if s.wav is None else np.concatenate([s.wav, x]).astype(np.float32)
        feats = fz.featurize(s.wav, sample_rate, inplace=True)
        s.wav = s.wav[160 * feats.shape[0]:]
        self.feed(sid, feats, is_end)

    def _audio_featurizer(self):
        if self._featurizer is None:
            from ..predict import AudioFeaturizer
            pre = self._preprocess_conf or {}
            pre = dict(pre) if isinstance(pre, dict) else vars(pre)
            pre.setdefault("n_mels", self.n_mels)
            self._featurizer = AudioFeaturizer(**pre)
        return self._featurizer

    def close(self, sid):
        s = self._sessions.pop(sid)
        self._free.append(s.slot)
        return {"text": s.text, "score": s.score}

    def pending(self):
        """Sessions that can take a step now: a full window, or (ended) at least `CONTEXT` buffered frames."""
        return [sid for sid, s in self._sessions.items()
                if s.feat.shape[0] >= WINDOW or (s.ended and s.feat.shape[0] >= CONTEXT)]

    def step(self):
        """One scheduling round: groups the ready sessions by window length (all full windows form one batch; short tails
        of ended sessions are batched among themselves) and steps each group once. Returns {sid: {'text', 'score'}}."""
        groups = {}
        for sid in self.pending():
            s = self._sessions[sid]
            t = WINDOW if s.feat.shape[0] >= WINDOW else s.feat.shape[0]
            groups.setdefault(t, []).append(sid)
        out = {}
        for t, sids in sorted(groups.items(), reverse=True):
            batch = np.stack([self._sessions[sid].feat[:t] for sid in sids])
            slots = [self._sessions[sid].slot for sid in sids]
            self.engine.sessions_step(batch, slots, self.required)
            probs = self.engine.ctc_probs(to_host=True)
            for b, sid in enumerate(sids):
                s = self._sessions[sid]
                s.score, s.text, s.ids, s.probs = greedy_decoder_chunk(probs_seq=probs[b], vocabulary=self.vocab,
                                                                         last_max_index_list=s.probs,
                                                                         last_max_prob_list=s.ids)
                # consumed: full windows advance by the stride and keep the 3 context frames; a tail is consumed entirely
                s.feat = s.feat[STRIDE:] if t == WINDOW else s.feat[:0]
                out[sid] = {"text": s.text, "score": s.score}
        return out
