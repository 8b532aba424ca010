"""generate_gestures - long-form inference driver around the hot path (reference: scripts/synthesize.py:36-209).

Same signature and result (direction vectors [n_total_frames, 27] as a float numpy array) for args.model in 'multimodal_context',
'joint_embedding' (decodes the speech latent, :132-133) and 'seq2seq' (one clip at a time: its word sequences are ragged, :134-136).
What is restructured for the device (the arithmetic per window is unchanged):
  * the audio slice and the frame-aligned word-index row of EVERY window are prepared on the host up front and copied in one H2D each;
  * the window chain (forward -> last n_pre_poses frames become the next window's seed poses, :122-126) runs without a host
    synchronisation: the reference copies every window to the host (:144) before it can launch the next one;
  * one D2H at the end; the 4-frame cross-fade between windows (:146-156) and the optional polynomial fade-out (:187-207) are tiny
    host post-processing steps on the collected [n_windows, n_poses, 27] array, bit-identical to the reference's NumPy code;
  * generate_gestures_batch runs the window chains of several clips in lock-step as one batch (the chain of one clip is inherently
    sequential; clips are independent)."""
import math
import random

import numpy as np
import torch


def get_words_in_time_range(word_list, start_time, end_time):
    """DataPreprocessor.get_words_in_time_range (scripts/data_loader/data_preprocessor.py:173-188)."""
    words = []
    for word in word_list:
        _, word_s, word_e = word[0], word[1], word[2]
        if word_s >= end_time:
            break
        if word_e <= start_time:
            continue
        words.append(word)
    return words


def _plan_windows(args, audio, audio_sr):
    """Window schedule of synthesize.py:58-68: (num_subdivision, unit_time, stride_time, audio_sample_length, clip_length)."""
    clip_length = len(audio) / audio_sr
    unit_time = args.n_poses / args.motion_resampling_framerate
    stride_time = (args.n_poses - args.n_pre_poses) / args.motion_resampling_framerate
    if clip_length < unit_time:
        num_subdivision = 1
    else:
        num_subdivision = math.ceil((clip_length - unit_time) / stride_time) + 1
    return num_subdivision, unit_time, stride_time, int(unit_time * audio_sr), clip_length


def _host_inputs(args, lang_model, audio, words, audio_sr):
    """Per-window audio slices [n_sub, audio_sample_length] float32 and frame-aligned word ids [n_sub, n_poses] int64 (:93-120)."""
    n_frames = args.n_poses
    n_sub, unit_time, stride_time, audio_len, clip_length = _plan_windows(args, audio, audio_sr)
    audio_rows = np.zeros((n_sub, audio_len), dtype=np.float32)
    text_rows = np.zeros((n_sub, n_frames), dtype=np.int64)
    end_padding = 0
    for i in range(n_sub):
        start_time = i * stride_time
        end_time = start_time + unit_time
        a0 = math.floor(start_time / clip_length * len(audio))
        seg = np.asarray(audio[a0:a0 + audio_len], dtype=np.float32)
        if len(seg) < audio_len and i == n_sub - 1:
            end_padding = audio_len - len(seg)
        audio_rows[i, :len(seg)] = seg
        frame_duration = (end_time - start_time) / n_frames
        for word in get_words_in_time_range(words, start_time, end_time):
            idx = max(0, int(np.floor((word[1] - start_time) / frame_duration)))
            text_rows[i, idx] = lang_model.get_word_index(word[0])
    return audio_rows, text_rows, n_sub, end_padding


def _crossfade_and_stack(windows, n_pre):
    """synthesize.py:146-161 on the collected windows [n_sub, n_poses, D] (float32, like the per-window .cpu().numpy() arrays)."""
    out_list = []
    for out_seq in windows:
        out_seq = out_seq.copy()
        if len(out_list) > 0:
            last_poses = out_list[-1][-n_pre:]
            out_list[-1] = out_list[-1][:-n_pre]
            for j in range(len(last_poses)):
                n = len(last_poses)
                prev = last_poses[j]
                nxt = out_seq[j]
                out_seq[j] = prev * (n - j) / (n + 1) + nxt * (j + 1) / (n + 1)
        out_list.append(out_seq)
    return np.vstack(out_list)


def _fade_out(out_dir_vec, args, end_padding_duration, audio_sr):
    """synthesize.py:187-207: fade to the mean pose with a weighted quadratic fit."""
    n_smooth = args.n_pre_poses
    start_frame = len(out_dir_vec) - int(end_padding_duration / audio_sr * args.motion_resampling_framerate)
    end_frame = start_frame + n_smooth * 2
    if len(out_dir_vec) < end_frame:
        out_dir_vec = np.pad(out_dir_vec, [(0, end_frame - len(out_dir_vec)), (0, 0)], mode='constant')
    out_dir_vec[end_frame - n_smooth:] = np.zeros((len(args.mean_dir_vec)))
    y = out_dir_vec[start_frame:end_frame]
    x = np.array(range(0, y.shape[0]))
    w = np.ones(len(y))
    w[0] = 5
    w[-1] = 5
    coeffs = np.polyfit(x, y, 2, w=w)
    fit_functions = [np.poly1d(coeffs[:, k]) for k in range(0, y.shape[1])]
    interpolated_y = np.transpose(np.asarray([fit_functions[k](x) for k in range(0, y.shape[1])]))
    out_dir_vec[start_frame:end_frame] = interpolated_y
    return out_dir_vec


def generate_gestures_batch(args, pose_decoder, lang_model, clips, audio_sr=16000, fade_out=False, device=None):
    """clips: list of dicts {'audio': 1-D float array, 'words': [[word, start, end], ...], 'vid': int or None, 'seed_seq': array or None}.
    Returns one [n_frames_i, D] array per clip.  All clips advance one window per forward call (batch = number of clips still running)."""
    assert args.model in ('multimodal_context', 'joint_embedding'), 'generate_gestures_batch: %r is not batched on the B200 path' % (args.model,)
    if device is None:
        device = next(pose_decoder.parameters()).device if hasattr(pose_decoder, 'parameters') else torch.device('cuda:0')
    D = len(args.mean_dir_vec)
    n_frames, n_pre = args.n_poses, args.n_pre_poses
    plans = [_host_inputs(args, lang_model, c['audio'], c['words'], audio_sr) for c in clips]
    order = sorted(range(len(clips)), key=lambda i: -plans[i][2])            # longest chain first: the running set is always a prefix
    n_max = plans[order[0]][2]
    C = len(clips)
    audio_len = plans[0][0].shape[1]
    audio_all = np.zeros((n_max, C, audio_len), dtype=np.float32)
    text_all = np.zeros((n_max, C, n_frames), dtype=np.int64)
    for slot, ci in enumerate(order):
        a, t, n_sub, _ = plans[ci]
        audio_all[:n_sub, slot] = a
        text_all[:n_sub, slot] = t
    audio_dev = torch.from_numpy(audio_all).to(device)
    text_dev = torch.from_numpy(text_all).to(device)
    vids = None
    if args.model == 'multimodal_context' and args.z_type == 'speaker':
        ids = []
        for ci in order:
            vid = clips[ci].get('vid')
            if not vid:
                vid = random.randrange(pose_decoder.z_obj.n_words)          # synthesize.py:72-73
            ids.append(vid)
        vids = torch.LongTensor(ids).to(device)
    pre_seq = torch.zeros((C, n_frames, D + 1))
    for slot, ci in enumerate(order):
        seed = clips[ci].get('seed_seq')
        if seed is not None:
            pre_seq[slot, 0:n_pre, :-1] = torch.Tensor(np.asarray(seed)[0:n_pre])
            pre_seq[slot, 0:n_pre, -1] = 1
    pre_seq = pre_seq.float().to(device)
    windows = torch.zeros((n_max, C, n_frames, D), device=device)
    with torch.no_grad():
        for i in range(n_max):
            live = sum(1 for ci in order if plans[ci][2] > i)               # clips whose chain still has window i (a prefix of `order`)
            if args.model == 'joint_embedding':                                # synthesize.py:132-133
                out = pose_decoder(text_dev[i, :live], audio_dev[i, :live], pre_seq[:live, 0:n_pre, :-1], None, 'speech')[6]
            else:
                out, *_ = pose_decoder(pre_seq[:live], text_dev[i, :live], audio_dev[i, :live], vids[:live] if vids is not None else None)
            windows[i, :live] = out
            pre_seq[:live, 0:n_pre, :-1] = out[:, -n_pre:]                  # seed hand-off (:122-126), stays on the device
            pre_seq[:live, 0:n_pre, -1] = 1
    win_host = windows.cpu().numpy()                                        # the only device->host copy
    results = [None] * C
    for slot, ci in enumerate(order):
        n_sub, end_padding = plans[ci][2], plans[ci][3]
        out_dir_vec = _crossfade_and_stack(win_host[:n_sub, slot], n_pre)
        if fade_out:
            out_dir_vec = _fade_out(out_dir_vec, args, end_padding, audio_sr)
        results[ci] = out_dir_vec
    return results


def _generate_seq2seq(args, pose_decoder, lang_model, audio, words, audio_sr, seed_seq, fade_out, device):
    """synthesize.py:134-136: the seq2seq baseline reads the window's word sequence [SOS, w1.., EOS] (ragged -> one clip at a time) and the
    seed poses; the chain still runs without a host synchronisation per window."""
    n_frames, n_pre, D = args.n_poses, args.n_pre_poses, len(args.mean_dir_vec)
    n_sub, unit_time, stride_time, _, _ = _plan_windows(args, audio, audio_sr)
    _, _, _, end_padding = _host_inputs(args, lang_model, audio, words, audio_sr)
    texts = []
    for i in range(n_sub):
        seq = get_words_in_time_range(words, i * stride_time, i * stride_time + unit_time)
        ids = [lang_model.SOS_token] + [lang_model.get_word_index(w[0]) for w in seq] + [lang_model.EOS_token]      # :108-117
        texts.append(torch.LongTensor(ids).unsqueeze(0).to(device))
    poses = torch.zeros((1, n_frames, D), device=device)                       # only the first n_pre frames are read (seq2seq_net.py:244-252)
    if seed_seq is not None:
        poses[0, 0:n_pre] = torch.as_tensor(np.asarray(seed_seq)[0:n_pre], dtype=torch.float32)
    windows = torch.zeros((n_sub, n_frames, D), device=device)
    with torch.no_grad():
        for i in range(n_sub):
            out = pose_decoder(texts[i], [texts[i].shape[1]], poses, None)
            windows[i] = out[0]
            poses[0, 0:n_pre] = out[0, -n_pre:]
    out_dir_vec = _crossfade_and_stack(windows.cpu().numpy(), n_pre)
    out_dir_vec = _smooth_window_joins(out_dir_vec, n_sub, n_frames, n_pre)
    return _fade_out(out_dir_vec, args, end_padding, audio_sr) if fade_out else out_dir_vec


def _smooth_window_joins(out_dir_vec, n_sub, n_frames, n_pre):
    """seq2seq only (synthesize.py:163-185): around the start of every window the stacked sequence is replaced, in place and window after
    window, by its own least-squares cubic over 3 * n_pre frames beginning n_pre frames before the join (2 * n_pre frames from frame 0
    for the first window).  The fit is unweighted (the reference builds end weights but never hands them to polyfit)."""
    stride = n_frames - n_pre
    for i in range(n_sub):
        lo = n_pre + i * stride - n_pre
        hi = lo + (3 if lo >= 0 else 2) * n_pre
        lo = max(lo, 0)
        seg = out_dir_vec[lo:hi]
        if len(seg) == 0:
            continue
        x = np.arange(len(seg), dtype=np.float64)
        coeffs = np.polyfit(x, seg, 3)                                          # one cubic per pose dimension: [4, D]
        out_dir_vec[lo:hi] = np.vander(x, 4) @ coeffs
    return out_dir_vec


def generate_gestures(args, pose_decoder, lang_model, audio, words, audio_sr=16000, vid=None, seed_seq=None, fade_out=False):
    """Drop-in for scripts/synthesize.py:36-209 (multimodal_context, joint_embedding, seq2seq)."""
    if args.model == 'seq2seq':
        return _generate_seq2seq(args, pose_decoder, lang_model, audio, words, audio_sr, seed_seq, fade_out,
                                 next(pose_decoder.parameters()).device)
    return generate_gestures_batch(args, pose_decoder, lang_model, [dict(audio=audio, words=words, vid=vid, seed_seq=seed_seq)],
                                   audio_sr=audio_sr, fade_out=fade_out)[0]
