"""Drop-in for the hot-path part of ppasr/predict.py::PPASRPredictor on a B200.

Mirrors `predict`, `predict_stream`, `reset_stream` and `decode` (predict.py:114-140,163-187,232-347): same
arguments, same `{'text', 'score'}` results, same window logic (67-frame window, stride 64, 3 cached
frames), same `None` returns when too few frames are buffered, same exceptions. The encoder + decode call
path runs in libppasr_b200.so; only the audio front end (outside the hot path) is host code:
`AudioFeaturizer` below restates audio_featurizer.py:37-69,120-138 with torchaudio's kaldi fbank, the CPU
twin of paddleaudio's. Inputs as in `_load_audio` (predict.py:142-161): ndarray (float, or integer scaled to [-1, 1)), path /
open file / bytes of a complete RIFF/WAVE PCM file (other containers need soundfile / PyAV: decode them first).

`predict_long` (predict.py:190-229) is mirrored around the VAD: the silero ONNX model itself is outside the hot path
(SURVEY §2 row 12), so the speech segments come from `speech_timestamps=` or from any object with the reference's
`get_speech_timestamps(samples, sample_rate)` assigned to `self.vad_predictor`; the segments of one file then go through
the encoder as ONE ragged batch. Not mirrored (out of scope): punctuation, ITN —
passing `use_pun=True` / `is_itn=True` raises.
"""
import numpy as np

from .decoders.ctc_greedy_decoder import greedy_decoder, greedy_decoder_chunk
from .infer_utils.inference_predictor import InferencePredictor, _get
from . import SUPPORT_MODEL
from .weights import read_vocab_file


class AudioFeaturizer:
    """ppasr/data_utils/featurizer/audio_featurizer.py:13-138 for feature_method='fbank' (the shipped default)."""

    def __init__(self, feature_method='fbank', n_mels=80, n_mfcc=40, sample_rate=16000, use_dB_normalization=True,
                 target_dB=-20, train=False):
        if feature_method != 'fbank':
            raise Exception('没有{}预处理方法'.format(feature_method))
        self._n_mels = n_mels
        self._target_sample_rate = sample_rate
        self._use_dB_normalization = use_dB_normalization
        self._target_dB = target_dB

    @staticmethod
    def _rms_db(samples):
        mean_square = np.mean(samples ** 2)  # audio.py:519-529 (float32 like the reference; silence -> 0 dB)
        if mean_square == 0:
            mean_square = 1
        return 10 * np.log10(mean_square)

    @staticmethod
    def to_float32(samples):
        """audio.py:24-32,532-546 (AudioSegment.__init__): integer arrays are scaled to [-1, 1), floats are cast, a [N, C]
        array is averaged over its channels. Always returns a fresh float32 array."""
        samples = np.asarray(samples)
        out = samples.astype('float32')
        if samples.dtype in (np.int8, np.int16, np.int32, np.int64):
            out *= (1. / 2 ** (np.iinfo(samples.dtype).bits - 1))
        elif samples.dtype.kind != 'f':
            raise TypeError("Unsupported sample type: %s." % samples.dtype)
        if out.ndim >= 2:
            out = np.mean(out, 1)
        return out

    @classmethod
    def from_pcm_bytes(cls, data, channels=1, samp_width=2):
        """audio.py:122-139 from_pcm_bytes -> buf_to_float (data_utils/utils.py:381-410): headerless little-endian PCM."""
        scale = 1.0 / float(1 << ((8 * samp_width) - 1))
        x = scale * np.frombuffer(data, "<i{:d}".format(samp_width)).astype(np.float32)
        if channels > 1:
            x = x.reshape(-1, channels)
        return cls.to_float32(x)

    def normalize_(self, samples):
        """audio.py:256-264,287-304 normalize() -> gain_db(): IN PLACE on the float32 buffer, like the reference (the scaled
        samples stay in the caller's buffer -- predict_stream relies on that, predict.py:262-274)."""
        gain = self._target_dB - self._rms_db(samples)
        if gain > 300.0:
            raise ValueError(f"无法将段规范化到{self._target_dB}dB，音频增益{gain}增益已经超过max_gain_db (300.0dB)")
        samples *= 10. ** (gain / 20.)
        return samples

    def featurize(self, samples, sample_rate=16000, inplace=False):
        """samples: float32 in [-1, 1]. Returns fbank [T, n_mels] float32. With inplace=True `samples` must be a float32 array
        and is left dB-normalised, as AudioFeaturizer.featurize leaves its AudioSegment (audio_featurizer.py:47-49)."""
        import torch
        import torchaudio
        if not inplace:
            samples = np.array(samples, dtype=np.float32)
        if sample_rate != self._target_sample_rate:
            raise Exception("resampling is outside the hot path: feed audio at %d Hz" % self._target_sample_rate)
        if self._use_dB_normalization:
            self.normalize_(samples)
        # audio.py:549-574 to('int16'): scale by 2^15 and clip
        s16 = np.clip(samples * 32768.0, -32768, 32767).astype(np.int16)
        wav = torch.from_numpy(s16.astype(np.float32)).unsqueeze(0)
        if wav.shape[1] < 400:
            return np.zeros((0, self._n_mels), dtype=np.float32)
        mat = torchaudio.compliance.kaldi.fbank(wav, num_mel_bins=self._n_mels, frame_length=25, frame_shift=10,
                                                dither=0.0, sample_frequency=float(self._target_sample_rate))
        return mat.numpy().astype(np.float32)


def _read_wav(src):
    """RIFF/WAVE integer PCM through the standard library (the reference reads files with soundfile / PyAV, predict.py:143-161,
    audio.py:56-71,110-119, which also decode FLAC, MP3, ...: those need a decoder and stay outside the hot path).
    Returns (float32 samples in [-1, 1) -- [N] or [N, channels] like soundfile.read(dtype='float32') --, sample_rate)."""
    import io
    import wave
    if isinstance(src, (bytes, bytearray)):
        src = io.BytesIO(bytes(src))
    try:
        with wave.open(src, 'rb') as w:
            nch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
    except (wave.Error, EOFError) as e:
        raise Exception(f'only RIFF/WAVE PCM files are decoded here ({e}); decode other formats to a NumPy array first')
    if width == 1:    # 8-bit WAV is unsigned
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        x = np.frombuffer(raw, dtype='<i2').astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0
    else:
        raise Exception(f'unsupported WAV sample width: {width} bytes')
    if nch > 1:
        x = x.reshape(-1, nch)
    return x, rate


class PPASRPredictor:
    # 7 fbank frames (25 ms window, 10 ms shift) = the shortest input Conv2dSubsampling4 turns into one encoder frame; a shorter
    # region yields no text here (in the reference the fbank / conv of such a region raises)
    MIN_SAMPLES = 400 + 6 * 160

    def __init__(self, configs=None, model_path='models/conformer_streaming_fbank/infer/', use_pun=False, use_gpu=True,
                 vocab_list=None, weights=None, device=0, model_tag=None, pun_model_dir=None, decoder_fallback=False):
        """configs: path of a reference yaml config (predict.py:36-41), or the loaded dict / attribute object, with the
        reference's keys: use_model, streaming, decoder, encoder_conf, preprocess_conf, dataset_conf.dataset_vocab,
        ctc_beam_search_decoder_conf. `vocab_list` / `weights` let tests pass in-memory objects instead of files. `model_tag`
        (download of a published model, predict.py:42-58) needs the network and is not supported."""
        if use_pun:
            raise Exception("punctuation restoration is outside the ppasr_b200 hot path")
        configs = self.load_configs(configs, model_tag)
        self.configs = configs
        self.use_model = _get(configs, 'use_model', 'conformer')
        assert self.use_model in SUPPORT_MODEL, f'没有该模型：{self.use_model}'   # predict.py:61
        self.streaming = bool(_get(configs, 'streaming', True))
        self.decoder = _get(configs, 'decoder', 'ctc_greedy')
        pre = _get(configs, 'preprocess_conf', {}) or {}
        pre = dict(pre) if isinstance(pre, dict) else vars(pre)
        self._audio_featurizer = AudioFeaturizer(**pre)
        if vocab_list is None:
            vocab_list = read_vocab_file(_get(_get(configs, 'dataset_conf', {}), 'dataset_vocab'))
        self.vocab_list = vocab_list
        self.running = False
        self.vad_predictor = None
        # streaming state (predict.py:69-72)
        self.remained_wav = None
        self.cached_feat = None
        self.greedy_last_max_prob_list = None
        self.greedy_last_max_index_list = None
        self._init_decoder(decoder_fallback)
        self.predictor = InferencePredictor(configs=configs, use_model=self.use_model, streaming=self.streaming,
                                            model_dir=model_path, use_gpu=use_gpu, weights=weights,
                                            vocab_size=len(vocab_list), device=device)

    # predict.py:92-105
    def _init_decoder(self, decoder_fallback=False):
        """ctc_beam_search: build the GPU beam-search decoder from `ctc_beam_search_decoder_conf` (the shipped beam_size 300 /
        cutoff_top_n 40 run on the GPU). A configuration outside what the GPU decoder implements (a KenLM binary or
        word-based language model, a missing LM file, beam_size > 512) RAISES UnsupportedDecoderConfig: silently decoding
        with another decoder would change the results of a stock config. Only with decoder_fallback=True does it degrade
        to ctc_greedy with a warning, as the reference does when its own decoder cannot be initialised (missing
        paddlespeech_ctcdecoders, predict.py:98-105)."""
        if self.decoder != 'ctc_beam_search':
            return
        from .decoders.beam_search_decoder import BeamSearchDecoder, UnsupportedDecoderConfig
        bconf = _get(self.configs, 'ctc_beam_search_decoder_conf', {}) or {}
        bconf = dict(bconf) if isinstance(bconf, dict) else vars(bconf)
        try:
            self.beam_search_decoder = BeamSearchDecoder(vocab_list=self.vocab_list, **bconf)
        except UnsupportedDecoderConfig as e:
            if not decoder_fallback:
                raise
            import warnings
            warnings.warn(f"ctc_beam_search is not available with this configuration ({e}); "
                          "【注意】现在已自动切换为ctc_greedy解码器，ctc_greedy解码器准确率相对较低。")
            self.decoder = 'ctc_greedy'

    @staticmethod
    def load_configs(configs, model_tag=None):
        """predict.py:36-60: a str is the path of a yaml file; dicts / attribute objects pass through."""
        if configs is None:
            raise Exception(f"no configs given: downloading the published model '{model_tag}' needs the network; pass the "
                            "yaml config of an exported model instead")
        if isinstance(configs, str):
            import yaml
            with open(configs, 'r', encoding='utf-8') as f:
                configs = yaml.load(f.read(), Loader=yaml.FullLoader)
        return configs

    # predict.py:114-140
    def decode(self, output_data, use_pun=False, is_itn=False):
        if use_pun or is_itn:
            raise Exception("punctuation / ITN are outside the ppasr_b200 hot path")
        if self.decoder == 'ctc_beam_search':
            result = self.beam_search_decoder.decode_beam_search_offline(probs_split=output_data)
        else:
            result = greedy_decoder(probs_seq=output_data, vocabulary=self.vocab_list)
        return result[0], result[1]

    # predict.py:142-161
    @staticmethod
    def _load_audio(audio_data, sample_rate=16000):
        """str path / open binary file / bytes of a complete WAV file / ndarray -> (samples, sample_rate)."""
        from io import BufferedReader
        if isinstance(audio_data, np.ndarray):
            return audio_data, sample_rate
        if isinstance(audio_data, str):
            import os
            assert os.path.exists(audio_data), f'文件不存在，请检查路径：{audio_data}'
            return _read_wav(audio_data)
        if isinstance(audio_data, (BufferedReader, bytes)):
            return _read_wav(audio_data)
        raise Exception(f'不支持该数据类型，当前数据类型为：{type(audio_data)}')

    # predict.py:163-187
    def predict(self, audio_data, use_pun=False, is_itn=False, sample_rate=16000):
        audio_data, sample_rate = self._load_audio(audio_data, sample_rate)
        samples = self._audio_featurizer.to_float32(audio_data)  # AudioSegment.__init__ (audio.py:24-32)
        audio_feature = self._audio_featurizer.featurize(samples, sample_rate, inplace=True)
        return self.predict_features(audio_feature, use_pun=use_pun, is_itn=is_itn)

    def predict_batch(self, audio_batch, n_samples=None):
        """Extension: a batch of 16 kHz waveforms float32 [B, N] -> [{'text', 'score'}], everything from the fbank
        front end (csrc/fbank.cu) to the greedy decode on the GPU; only the decoded ids come back to the host."""
        from .featurizer import GpuFbank
        if self.decoder != 'ctc_greedy':
            raise Exception("predict_batch implements the ctc_greedy decoder")
        if getattr(self, '_gpu_fbank', None) is None:
            pre = _get(self.configs, 'preprocess_conf', {}) or {}
            pre = dict(pre) if isinstance(pre, dict) else vars(pre)
            self._gpu_fbank = GpuFbank(n_mels=pre.get('n_mels', 80), sample_rate=pre.get('sample_rate', 16000),
                                       use_dB_normalization=pre.get('use_dB_normalization', True),
                                       target_dB=pre.get('target_dB', -20))
        feats, counts = self._gpu_fbank.featurize_batch(audio_batch, n_samples)
        if n_samples is None:
            res = self.predictor.predict_decode(feats, np.asarray(counts, dtype=np.int64), vocabulary=self.vocab_list)
            return [{'text': t, 'score': s} for s, t in res]
        # ragged batch: every utterance is decoded over exactly the frames a run of its own would produce
        # (out_frames(its fbank frames), not the 1-2 extra frames the batch mask keeps, subsampling.py:115), from the per-frame
        # arg-max ids / probabilities of the fused CTC head
        from .decoders.ctc_greedy_decoder import collapse_frames
        eng = self.predictor.engine
        # a padded batch marks frame j valid iff 4 j < len (subsampling.py:115), 1-2 frames more than the utterance has on its
        # own; lengths of the form 4 T' - 3 make the batch mask end exactly at T' = ((frames - 1) // 2 - 1) // 2, so that
        # attention and the conv masks see what a stand-alone run sees
        t_own = [max(((int(c) - 1) // 2 - 1) // 2, 0) for c in counts]
        lens = counts if self.use_model == 'deepspeech2' else [4 * t - 3 if t > 0 else 0 for t in t_own]  # DS2: conv.py:20 rule
        eng.encode(feats, np.asarray(lens, dtype=np.int64))
        _, _, _, frame_ids, frame_probs = eng.ctc_greedy(to_host=True, with_frames=True)
        out = []
        for b, c in enumerate(counts):
            n = int(eng.lib.ppasr_b200_out_frames(eng._ctx, int(c))) if c > 0 else 0
            score, text = collapse_frames(frame_ids[b], frame_probs[b], self.vocab_list, n_frames=max(n, 0))
            out.append({'text': text, 'score': score})
        return out

    # predict.py:190-229
    def predict_long(self, audio_data, use_pun=False, is_itn=False, sample_rate=16000, speech_timestamps=None, batched=True,
                     max_batch_samples=32 * 30 * 16000):
        """Long-form recognition: the speech regions of one recording -> one result. Same composition as the reference: every
        region is recognised on its own (own dB normalisation), non-empty texts are joined with '，', the score is the mean of
        the region scores rounded to 2 places. `speech_timestamps`: [{'start': s, 'end': e}, ...] in samples, what
        VADPredictor.get_speech_timestamps returns (vad_predictor.py:106-175); when None, `self.vad_predictor` is asked.
        batched=True (ctc_greedy only) sends the regions through `predict_batch` in ragged batches of at most
        `max_batch_samples` padded samples instead of one call per region."""
        if use_pun or is_itn:
            raise Exception("punctuation / ITN are outside the ppasr_b200 hot path")
        audio_data, sample_rate = self._load_audio(audio_data, sample_rate)
        samples = self._audio_featurizer.to_float32(audio_data)
        if sample_rate != self._audio_featurizer._target_sample_rate:
            raise Exception("resampling is outside the hot path: feed audio at %d Hz" % self._audio_featurizer._target_sample_rate)
        if speech_timestamps is None:
            if self.vad_predictor is None:
                raise Exception("predict_long needs speech_timestamps= or a vad_predictor with get_speech_timestamps()")
            speech_timestamps = self.vad_predictor.get_speech_timestamps(samples, sample_rate)
        segments = [samples[int(t['start']): int(t['end'])] for t in speech_timestamps]
        if batched and self.decoder == 'ctc_greedy':
            results = self._predict_segments_batched(segments, max_batch_samples)
        else:
            results = [self.predict(audio_data=seg, sample_rate=sample_rate) if len(seg) >= self.MIN_SAMPLES
                       else {'text': '', 'score': 0} for seg in segments]
        texts, scores = '', []
        for result in results:
            score, text = result['score'], result['text']
            if text != '':
                texts = texts + '，' + text
            scores.append(score)
        if texts[:1] == '，':
            texts = texts[1:]
        # the reference divides by len(scores) and indexes texts[0] unguarded (predict.py:219,226): no region -> exception there,
        # an empty result here
        return {'text': texts, 'score': round(sum(scores) / len(scores), 2) if scores else 0}

    def _predict_segments_batched(self, segments, max_batch_samples):
        """Regions -> ragged batches for predict_batch, in the original order; a batch is closed when its padded size
        (count x longest region) would exceed max_batch_samples."""
        results = [None] * len(segments)
        group = []

        def flush():
            if not group:
                return
            n = np.array([len(segments[i]) for i in group], dtype=np.int32)
            batch = np.zeros((len(group), int(n.max())), dtype=np.float32)
            for r, i in enumerate(group):
                batch[r, :n[r]] = segments[i]
            for i, res in zip(group, self.predict_batch(batch, n_samples=n)):
                results[i] = res
            group.clear()

        longest = 0
        for i, seg in enumerate(segments):
            if len(seg) < self.MIN_SAMPLES:
                results[i] = {'text': '', 'score': 0}
                continue
            if group and (len(group) + 1) * max(longest, len(seg)) > max_batch_samples:
                flush()
                longest = 0
            group.append(i)
            longest = max(longest, len(seg))
        flush()
        return results

    def predict_features(self, audio_feature, use_pun=False, is_itn=False):
        """Same as predict() from the featurizer output on ([T, n_mels] fp32)."""
        input_data = np.array(audio_feature).astype(np.float32)[np.newaxis, :]
        audio_len = np.array([input_data.shape[1]]).astype(np.int64)
        output_data = self.predictor.predict(input_data, audio_len)[0]
        score, text = self.decode(output_data=output_data, use_pun=use_pun, is_itn=is_itn)
        return {'text': text, 'score': score}

    # predict.py:232-337
    def predict_stream(self, audio_data, is_end=False, use_pun=False, is_itn=False, channels=1, samp_width=2,
                       sample_rate=16000):
        if not self.streaming:
            raise Exception(f"不支持改该模型流式识别，当前模型：{self.use_model}")
        if isinstance(audio_data, np.ndarray):    # predict.py:253-254 AudioSegment.from_ndarray
            audio_data = self._audio_featurizer.to_float32(audio_data)
        elif isinstance(audio_data, bytes):       # predict.py:255-257
            audio_data = self._audio_featurizer.from_pcm_bytes(audio_data, channels=channels, samp_width=samp_width)
        else:
            raise Exception(f'不支持该数据类型，当前数据类型为：{type(audio_data)}')
        if self.remained_wav is None:
            self.remained_wav = audio_data
        else:
            self.remained_wav = np.concatenate([self.remained_wav, audio_data]).astype(np.float32)
        # featurise ALL remaining audio -- which dB-normalises it IN PLACE, so the unconsumed tail carried to the next call
        # stays scaled exactly as in the reference (predict.py:262-274) -- then drop the consumed samples
        x_chunk = self._audio_featurizer.featurize(self.remained_wav, sample_rate, inplace=True)
        x_chunk = np.array(x_chunk).astype(np.float32)[np.newaxis, :]
        self.remained_wav = self.remained_wav[160 * x_chunk.shape[1]:]
        return self.predict_stream_features(x_chunk, is_end=is_end, use_pun=use_pun, is_itn=is_itn)

    def predict_stream_features(self, x_chunk, is_end=False, use_pun=False, is_itn=False):
        """predict.py:269-337 from the featurizer output on: x_chunk [1, n, n_mels] new frames."""
        if use_pun or is_itn:
            raise Exception("punctuation / ITN are outside the ppasr_b200 hot path")
        if self.cached_feat is None:
            self.cached_feat = x_chunk
        else:
            self.cached_feat = np.concatenate([self.cached_feat, x_chunk], axis=1)
        decoding_chunk_size = 16
        context = 7
        subsampling = 4
        cached_feature_num = context - subsampling
        decoding_window = (decoding_chunk_size - 1) * subsampling + context
        stride = subsampling * decoding_chunk_size
        num_frames = self.cached_feat.shape[1]
        if num_frames < decoding_window and not is_end:
            return None
        if num_frames < context:
            return None
        left_frames = context if is_end else decoding_window
        score, text, end = None, None, None
        for cur in range(0, num_frames - left_frames + 1, stride):
            end = min(cur + decoding_window, num_frames)
            x = self.cached_feat[:, cur:end, :]
            if self.use_model == 'deepspeech2':  # predict.py:302-303
                output_chunk_probs, output_lens = self.predictor.predict_chunk_deepspeech(x_chunk=x)
            elif 'former' in self.use_model:     # predict.py:304-309
                num_decoding_left_chunks = -1
                required_cache_size = decoding_chunk_size * num_decoding_left_chunks
                output_chunk_probs = self.predictor.predict_chunk_conformer(x_chunk=x, required_cache_size=required_cache_size)
                output_lens = np.array([output_chunk_probs.shape[1]])
            else:
                raise Exception(f'当前模型不支持该方法，当前模型为：{self.use_model}')
            if self.decoder == 'ctc_beam_search':
                score, text = self.beam_search_decoder.decode_chunk(probs=output_chunk_probs, logits_lens=output_lens)
            else:
                score, text, self.greedy_last_max_prob_list, self.greedy_last_max_index_list = \
                    greedy_decoder_chunk(probs_seq=output_chunk_probs[0], vocabulary=self.vocab_list,
                                         last_max_index_list=self.greedy_last_max_index_list,
                                         last_max_prob_list=self.greedy_last_max_prob_list)
        self.cached_feat = self.cached_feat[:, end - cached_feature_num:, :]
        return {'text': text, 'score': score}

    # predict.py:340-347
    def reset_stream(self):
        self.predictor.reset_stream()
        self.remained_wav = None
        self.cached_feat = None
        self.greedy_last_max_prob_list = None
        self.greedy_last_max_index_list = None
        if self.decoder == 'ctc_beam_search':
            self.beam_search_decoder.reset_decoder()
