"""Text -> waveform in one call: the body of the reference's inference loop (inference.py:84-97) on the B200 modules.

    y_enc, y_dec, attn = generator.forward(x, x_lengths, n_timesteps, temperature=1.5, stoc=False, spk=spk, length_scale=1)
    audio = (vocoder.forward(y_dec).cpu().squeeze().clamp(-1, 1).numpy() * 32768).astype(np.int16)

`generator` is `GradTTS` (text encoder + MAS-free inference alignment + reverse-diffusion decoder), `vocoder` the HiFi-GAN
`Generator`; everything between the token ids and the int16 samples runs on the device.  Text normalisation / phonemisation
(text/, cmudict) and file I/O stay with the caller, as in the reference script.
"""
import torch


@torch.no_grad()
def synthesize(generator, vocoder, x, x_lengths, n_timesteps=10, temperature=1.5, spk=None, length_scale=1.0, to_int16=True):
    """Returns (audio, y_dec, attn): audio is (B, samples) int16 on the host when `to_int16` (as the reference writes it), else the
    (B, 1, samples) float waveform on the device.  Padded frames of shorter utterances are vocoded too, as in the reference."""
    y_enc, y_dec, attn = generator.forward(x, x_lengths, n_timesteps=n_timesteps, temperature=temperature, stoc=False, spk=spk,
                                           length_scale=length_scale)
    wav = vocoder.forward(y_dec)
    if not to_int16:
        return wav, y_dec, attn
    audio = (wav.squeeze(1).clamp(-1, 1) * 32768).to(torch.int16).cpu()        # .astype(np.int16) truncates toward zero, like .to(int16)
    return audio, y_dec, attn
