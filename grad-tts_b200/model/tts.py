"""GradTTS glue with the reference's signature (reference model/tts.py:21-108).

Only the boundary is reproduced here: speaker embedding lookup, encoder call, duration -> length -> mask ->
`generate_path` -> `mu_y`, the single RNG draw `z = mu_y + randn_like(mu_y) / temperature` (kept in PyTorch so
the same seed gives the same z as the reference) and the decoder call, which runs on the sm_100a
kernels.  The text encoder defaults to the native one (model/text_encoder.py here, csrc/text_encoder.cu; inference only); for
training pass `encoder=` (any module with the reference `TextEncoder.forward(x, x_lengths, spk) -> mu_x, logw, x_mask` contract,
e.g. the reference's own PyTorch class).
"""
import math
import random

import torch

from . import align, monotonic_align
from .base import BaseModule
from .diffusion import Diffusion
from .text_encoder import TextEncoder
from .utils import sequence_mask, generate_path, duration_loss, fix_len_compatibility  # noqa: F401


class GradTTS(BaseModule):
    def __init__(self, n_vocab, n_spks, spk_emb_dim, n_enc_channels, filter_channels, filter_channels_dp,
                 n_heads, n_enc_layers, enc_kernel, enc_dropout, window_size,
                 n_feats, dec_dim, beta_min, beta_max, pe_scale, encoder=None):
        super().__init__()
        self.n_vocab = n_vocab
        self.n_spks = n_spks
        self.spk_emb_dim = spk_emb_dim
        self.n_enc_channels = n_enc_channels
        self.filter_channels = filter_channels
        self.filter_channels_dp = filter_channels_dp
        self.n_heads = n_heads
        self.n_enc_layers = n_enc_layers
        self.enc_kernel = enc_kernel
        self.enc_dropout = enc_dropout
        self.window_size = window_size
        self.n_feats = n_feats
        self.dec_dim = dec_dim
        self.beta_min = beta_min
        self.beta_max = beta_max
        self.pe_scale = pe_scale

        if self.n_spks == -1:                                # model/tts.py:43-47
            self.spk_emb = None
        elif self.n_spks > 1:
            self.spk_emb = torch.nn.Embedding(n_spks, spk_emb_dim)
        if encoder is None:
            # as the reference builds it (model/tts.py:49-51): without spk_emb_dim / n_spks, i.e. the encoder ignores the speaker
            encoder = TextEncoder(n_vocab, n_feats, n_enc_channels, filter_channels, filter_channels_dp,
                                  n_heads, n_enc_layers, enc_kernel, enc_dropout, window_size)
        self.encoder = encoder
        self.decoder = Diffusion(n_feats, dec_dim, n_spks, spk_emb_dim, beta_min, beta_max, pe_scale)

    @torch.no_grad()
    def forward(self, x, x_lengths, n_timesteps, temperature=1.0, stoc=False, spk=None, length_scale=1.0):
        """Text -> (encoder_outputs, decoder_outputs, attn); reference model/tts.py:54-108."""
        x, x_lengths = self.relocate_input([x, x_lengths])
        if self.n_spks > 1:
            spk = self.spk_emb(spk)                          # :77-79 (n_spks == -1 passes spk through)

        mu_x, logw, x_mask = self.encoder(x, x_lengths, spk)               # :84

        w = torch.exp(logw) * x_mask                                        # :86-90
        w_ceil = torch.ceil(w) * length_scale
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_max_length = int(y_lengths.max())
        y_max_length_ = fix_len_compatibility(y_max_length)

        y_mask = sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)    # :93-95
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)

        mu_y = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2))        # :98-100
        mu_y = mu_y.transpose(1, 2)
        encoder_outputs = mu_y[:, :, :y_max_length]

        z = mu_y + torch.randn_like(mu_y, device=mu_y.device) / temperature              # :103
        decoder_outputs = self.decoder(z, y_mask, mu_y, n_timesteps, stoc, spk)           # :105  (sm_100a kernels)
        decoder_outputs = decoder_outputs[:, :, :y_max_length]

        return encoder_outputs, decoder_outputs, attn[:, :, :y_max_length]               # :108 (slices the text axis)

    @torch.no_grad()
    def align(self, mu_x, x_mask, y, y_mask):
        """MAS between encoder outputs and a mel (reference model/tts.py:139-152, 224-231): the
        log-prior kernel (csrc/align.cu) followed by `monotonic_align.maximum_path`, both on the device."""
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        log_prior = align.log_prior(mu_x, y)                                             # :143-149, one kernel
        attn = monotonic_align.maximum_path(log_prior, attn_mask.squeeze(1))
        return attn.detach()

    @torch.no_grad()
    def align_outputs(self, attn, mu_x, x_mask):
        """(logw_, mu_y) derived from a MAS path (reference model/tts.py:155, 184-185)."""
        return align.logw_from_path(attn, x_mask), align.mu_y_from_path(attn, mu_x)

    def get_score_model(self, x, x_lengths, y, y_lengths, spk=None):
        """Score model for a speech/text pair (reference model/tts.py:197-254): encoder + device MAS +
        closure over the sm_100a estimator forward."""
        x, x_lengths, y, y_lengths = self.relocate_input([x, x_lengths, y, y_lengths])
        if self.n_spks > 1:
            spk = self.spk_emb(spk)
        mu_x, logw, x_mask = self.encoder(x, x_lengths, spk)
        y_max_length = y.shape[-1]
        y_mask = sequence_mask(y_lengths, y_max_length).unsqueeze(1).to(x_mask)
        attn = self.align(mu_x, x_mask, y, y_mask)
        mu_y = align.mu_y_from_path(attn, mu_x)                                          # :184-185 / :233-234
        estimator = self.decoder.estimator

        class ScoreModel(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.y_mask, self.mu_y, self.spk = y_mask, mu_y, spk
                self.estimator = estimator

            def forward(self, x, t):
                return self.estimator(x=x, mask=self.y_mask, mu=self.mu_y, t=t, spk=self.spk)

        return ScoreModel(), mu_y, spk, y_mask

    def compute_loss(self, x, x_lengths, y, y_lengths, spk=None, out_size=None):
        """The three training losses (reference model/tts.py:110-194): duration loss against the MAS durations, prior loss,
        diffusion loss.  The alignment stage (log-prior, MAS), the estimator forward and its backward run as sm_100a kernels.
        In training mode with gradients enabled the losses carry the autograd graph (diffusion loss -> estimator parameters, mu_y ->
        encoder; duration / prior losses -> encoder), so `sum(losses).backward()` trains the model like the reference's train.py."""
        train = torch.is_grad_enabled() and self.training
        x, x_lengths, y, y_lengths = self.relocate_input([x, x_lengths, y, y_lengths])
        if self.n_spks > 1:
            spk = self.spk_emb(spk)                                                       # :130-136
        mu_x, logw, x_mask = self.encoder(x, x_lengths, spk)                              # :139
        y_max_length = y.shape[-1]
        y_mask = sequence_mask(y_lengths, y_max_length).unsqueeze(1).to(x_mask)
        attn = self.align(mu_x, x_mask, y, y_mask)                                        # :146-152
        logw_ = align.logw_from_path(attn, x_mask)                                        # :155
        dur_loss = duration_loss(logw, logw_, x_lengths)

        if out_size is not None:                                                          # :159-181, random segment per item
            max_offset = (y_lengths - out_size).clamp(0).tolist()
            offsets = [random.choice(range(0, end)) if end > 0 else 0 for end in max_offset]
            attn_cut = torch.zeros(attn.shape[0], attn.shape[1], out_size, dtype=attn.dtype, device=attn.device)
            y_cut = torch.zeros(y.shape[0], self.n_feats, out_size, dtype=y.dtype, device=y.device)
            cut_lengths = []
            for i, lo in enumerate(offsets):
                n = out_size + min(int(y_lengths[i]) - out_size, 0)
                cut_lengths.append(n)
                y_cut[i, :, :n] = y[i, :, lo:lo + n]
                attn_cut[i, :, :n] = attn[i, :, lo:lo + n]
            y_mask = sequence_mask(torch.tensor(cut_lengths, device=y.device)).unsqueeze(1).to(y_mask)
            if y_mask.shape[-1] < out_size:
                y_mask = torch.nn.functional.pad(y_mask, (0, out_size - y_mask.shape[-1]))
            attn, y = attn_cut, y_cut

        if train:                                                                         # :184-185 (autograd into the encoder)
            mu_y = torch.matmul(attn.transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
        else:
            mu_y = align.mu_y_from_path(attn, mu_x)
        diff_loss, _ = self.decoder.compute_loss(y, y_mask, mu_y, spk)                    # :188
        prior_loss = torch.sum(0.5 * ((y - mu_y) ** 2 + math.log(2 * math.pi)) * y_mask)  # :191-192
        prior_loss = prior_loss / (torch.sum(y_mask) * self.n_feats)
        return dur_loss, prior_loss, diff_loss
