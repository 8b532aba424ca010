"""init_model / train_epoch - the model switch and the per-batch dispatch of scripts/train.py:36-68,166-226 around the step functions.

`init_model` is the reference's constructor switch for all five model families (multimodal_context, joint_embedding, gesture_autoencoder,
seq2seq, speech2gesture).  `train_epoch` is the inner loop of train_epochs
(train.py:166-226): batches come through train_eval.staging.DevicePrefetcher (the copy of batch i+1 overlaps step i; the reference does a
blocking `.to(device)` per tensor, and copies the unused spectrogram too, :171-176), speaker ids are looked up like :178-183, the step
function is chosen by args.model like :186-203, and the loss meters are fed like :205-209.  Tensorboard, checkpoints, sample videos and
the epoch-level bookkeeping of train_epochs (:70-164,211-231) are the caller's: this module has no file or logging side effects."""
import torch

from model import vocab
from model.embedding_net import EmbeddingNet
from model.multimodal_context_net import ConvDiscriminator, PoseGenerator
from model.seq2seq_net import Seq2SeqNet
from model import speech2gesture
from tgb200 import _lib
from train_eval.train_gan import train_iter_gan
from train_eval.train_joint_embed import train_iter_embed
from train_eval.train_seq2seq import train_iter_seq2seq
from train_eval.train_speech2gesture import train_iter_speech2gesture

LOSS_NAMES = ('loss', 'var_loss', 'gen', 'dis', 'KLD', 'DIV_REG')           # train.py:72-73


def init_model(args, lang_model, speaker_model, pose_dim, _device):
    """train.py:36-68 -> (generator, discriminator, loss_fn)."""
    n_frames = args.n_poses
    generator = discriminator = loss_fn = None
    if args.model == 'multimodal_context':
        generator = PoseGenerator(args, n_words=lang_model.n_words, word_embed_size=args.wordembed_dim,
                                  word_embeddings=lang_model.word_embedding_weights, z_obj=speaker_model, pose_dim=pose_dim).to(_device)
        discriminator = ConvDiscriminator(pose_dim).to(_device)
    elif args.model == 'joint_embedding':
        generator = EmbeddingNet(args, pose_dim, n_frames, lang_model.n_words, args.wordembed_dim, lang_model.word_embedding_weights,
                                 mode='random').to(_device)
    elif args.model == 'gesture_autoencoder':
        generator = EmbeddingNet(args, pose_dim, n_frames, lang_model.n_words, args.wordembed_dim, lang_model.word_embedding_weights,
                                 mode='pose').to(_device)
    elif args.model == 'seq2seq':
        generator = Seq2SeqNet(args, pose_dim, n_frames, lang_model.n_words, args.wordembed_dim, lang_model.word_embedding_weights).to(_device)
        loss_fn = torch.nn.L1Loss()
    elif args.model == 'speech2gesture':                                                                 # train.py:59-62
        generator = speech2gesture.Generator(n_frames, pose_dim, args.n_pre_poses).to(_device)
        discriminator = speech2gesture.Discriminator(pose_dim).to(_device)
        loss_fn = torch.nn.L1Loss()
    else:
        raise NotImplementedError('unknown model %r' % (args.model,))
    return generator, discriminator, loss_fn


class Meter:
    """utils/average_meter.py: running sample-weighted average."""

    def __init__(self, name):
        self.name, self.sum, self.count = name, 0.0, 0

    def update(self, val, n=1):
        self.sum += val * n
        self.count += n

    @property
    def avg(self):
        return self.sum / self.count if self.count else 0.0


def train_epoch(args, epoch, train_data_loader, generator, discriminator, gen_optimizer, dis_optimizer, speaker_model=None, device=None,
                on_step=None):
    """One pass over train_data_loader (train.py:166-226).  Batches are the reference's collate tuples
    (in_text, text_lengths, in_text_padded, _, target_vec, in_audio, in_spec, aux_info).  on_step(iter_idx, loss_dict, batch_size) is
    called after every step (tensorboard / printing hook).  Returns {name: sample-weighted average} for the losses that occurred."""
    if device is None:
        device = next(generator.parameters()).device
    device = torch.device(device)
    meters = {n: Meter(n) for n in LOSS_NAMES}

    def strip(data):                      # the spectrogram is read by speech2gesture only: do not ship it otherwise (train.py:175)
        in_text, text_lengths, in_text_padded, _, target_vec, in_audio, in_spec, aux_info = data
        return in_text, text_lengths, in_text_padded, None, target_vec, in_audio, in_spec if args.model == 'speech2gesture' else None, aux_info

    batches = (strip(d) for d in train_data_loader)
    if device.type == 'cuda' and not _lib.TRACE_ONLY:
        from train_eval.staging import DevicePrefetcher
        batches = DevicePrefetcher(batches, device)
    for iter_idx, data in enumerate(batches):
        in_text, text_lengths, in_text_padded, _, target_vec, in_audio, in_spec, aux_info = data
        batch_size = target_vec.size(0)
        vid_indices = []
        if speaker_model and isinstance(speaker_model, vocab.Vocab):                                    # :178-183
            vid_indices = torch.LongTensor([speaker_model.word2index[vid] for vid in aux_info['vid']]).to(device)
        if args.model == 'multimodal_context':
            loss = train_iter_gan(args, epoch, in_text_padded, in_audio, target_vec, vid_indices if len(vid_indices) else None,
                                  generator, discriminator, gen_optimizer, dis_optimizer)
        elif args.model == 'joint_embedding':
            loss = train_iter_embed(args, epoch, in_text_padded, in_audio, target_vec, generator, gen_optimizer, mode='random')
        elif args.model == 'gesture_autoencoder':
            loss = train_iter_embed(args, epoch, in_text_padded, in_audio, target_vec, generator, gen_optimizer)
        elif args.model == 'seq2seq':
            lengths = text_lengths.cpu() if torch.is_tensor(text_lengths) else text_lengths              # pack_padded_sequence wants host lengths
            loss = train_iter_seq2seq(args, epoch, in_text, lengths, target_vec, generator, gen_optimizer)
        elif args.model == 'speech2gesture':                                                             # :199-201
            loss = train_iter_speech2gesture(args, in_spec, target_vec, generator, discriminator, gen_optimizer, dis_optimizer, None)
        else:
            raise NotImplementedError(args.model)
        for name, val in loss.items():                                                                   # :205-209
            if name in meters:
                meters[name].update(val, batch_size)
        if on_step is not None:
            on_step(iter_idx, loss, batch_size)
    return {n: m.avg for n, m in meters.items() if m.count > 0}
