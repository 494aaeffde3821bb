"""where the wall-clock time of one text -> waveform call goes (host side included)"""
import importlib
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("grad-tts_b200")
dev = torch.device("cuda:0")
ecfg = pkg.synth.TEXT_ENCODER_CONFIGS["ref"]
net = pkg.GradTTS(ecfg["n_vocab"], 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000)
net.encoder.load_state_dict(pkg.synth.make_text_encoder_state_dict(ecfg, seed=1))
net.decoder.load_state_dict(pkg.synth.make_decoder_state_dict(1, seed=0, g=0.05))
net = net.to(dev).eval()
vcfg = pkg.synth.VOCODER_CONFIGS["v1"]
voc = pkg.hifigan.Generator(pkg.hifigan.AttrDict(vcfg))
voc.load_state_dict(pkg.synth.make_vocoder_state_dict(vcfg, seed=1))
voc = voc.to(dev).eval()
voc.remove_weight_norm()
x, lengths, _ = pkg.synth.make_text_inputs(ecfg, 1, 100, seed=3, ragged=False)


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    torch.manual_seed(0)
    t0 = t()
    xd, ld = x.to(dev), lengths.to(dev)
    t1 = t()
    with torch.no_grad():
        mu_x, logw, x_mask = net.encoder(xd, ld, None)
        t2 = t()
        utils = importlib.import_module("grad-tts_b200.model.utils")
        w = torch.exp(logw) * x_mask
        w_ceil = torch.ceil(w)
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_max_length = int(y_lengths.max())
        y_max_length_ = utils.fix_len_compatibility(y_max_length)
        y_mask = utils.sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn = utils.generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
        mu_y = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
        z = mu_y + torch.randn_like(mu_y) / 1.5
        t3 = t()
        y = net.decoder(z, y_mask, mu_y, 10, False, None)
        t4 = t()
        wav = voc(y[:, :, :y_max_length])
        t5 = t()
        audio = (wav.squeeze(1).clamp(-1, 1) * 32768).to(torch.int16).cpu()
        t6 = t()
    print(f"rep {rep}: h2d {1e3*(t1-t0):.3f} encoder {1e3*(t2-t1):.3f} glue {1e3*(t3-t2):.3f} decoder {1e3*(t4-t3):.3f} "
          f"vocoder {1e3*(t5-t4):.3f} to_int16+d2h {1e3*(t6-t5):.3f} total {1e3*(t6-t0):.3f} ms (frames {y_max_length})", flush=True)
