"""What the reference's UNCHANGED train.py does per mini-batch, re-typed for the tests (the GPU box has no
/root/reference): the call sequence of train.py:528-580 with its inline torch BCE expressions,
torch.nn.utils.clip_grad_norm_ and torch.optim.Adagrad, written against the public `gantts` surface only
(`gantts.multistream`, `gantts.seqloss`, model callables).  Used to check the DROP-IN mode: the reference's
own step logic running on top of the B200 modules must reproduce the golden vectors the reference produced
on top of its own modules."""
import torch

from gantts.multistream import get_static_features, get_static_stream_sizes, multi_stream_mlpg, select_streams
from gantts.seqloss import MaskedMSELoss, sequence_mask


def selected_static_stream(y_static, hp):                      # train.py:232-242
    sizes = get_static_stream_sizes(hp.stream_sizes, hp.has_dynamic_features, len(hp.windows))
    sel = select_streams(y_static, sizes, streams=hp.adversarial_streams)
    if hp.mask_nth_mgc_for_adv_loss > 0:
        sel = sel[:, :, hp.mask_nth_mgc_for_adv_loss:]
    return sel


def train_step(model_g, model_d, opt_g, opt_d, x, y, lengths, R, hp, adv_w=1.0, mse_w=0.0, mge_w=1.0, eps=1e-20):
    y_static = get_static_features(y, len(hp.windows), hp.stream_sizes, hp.has_dynamic_features)   # :528
    mask = sequence_mask(lengths).unsqueeze(-1)                                                     # :535
    opt_g.zero_grad()
    opt_d.zero_grad()
    y_hat = model_g(x, lengths=lengths)                                                             # :345
    y_hat_static = multi_stream_mlpg(y_hat, R, hp.stream_sizes, hp.has_dynamic_features)           # :352
    # update_discriminator, train.py:245-279
    real_in, fake_in = selected_static_stream(y_static, hp), selected_static_stream(y_hat_static, hp)
    if hp.discriminator_linguistic_condition:
        real_in, fake_in = torch.cat((x, real_in), -1), torch.cat((x, fake_in), -1)
    T = mask.sum().item()
    D_real = model_d(real_in, lengths=lengths)
    real_correct = ((D_real > 0.5).float() * mask).sum().item()
    D_fake = model_d(fake_in, lengths=lengths)
    fake_correct = ((D_fake < 0.5).float() * mask).sum().item()
    loss_real_d = -(torch.log(D_real + eps) * mask).sum() / T
    loss_fake_d = -(torch.log(1 - D_fake + eps) * mask).sum() / T
    loss_d = loss_real_d + loss_fake_d
    loss_d.backward(retain_graph=True)
    torch.nn.utils.clip_grad_norm_(model_d.parameters(), 1.0)
    opt_d.step()
    # update_generator, train.py:282-320
    crit = MaskedMSELoss()
    loss_mge = crit(y_hat_static, y_static, mask=mask)
    loss_mse = crit(y_hat, y, mask=mask)
    fake_in = selected_static_stream(y_hat_static, hp)
    if hp.discriminator_linguistic_condition:
        fake_in = torch.cat((x, fake_in), -1)
    loss_adv = -(torch.log(model_d(fake_in, lengths=lengths) + eps) * mask).sum() / T
    loss_g = (mse_w * loss_mse + mge_w * loss_mge) + adv_w * loss_adv
    loss_g.backward()
    torch.nn.utils.clip_grad_norm_(model_g.parameters(), 1.0)
    opt_g.step()
    return ([loss_d.item(), loss_fake_d.item(), loss_real_d.item(), loss_mse.item(), loss_mge.item(),
             loss_adv.item(), loss_g.item()], [real_correct, fake_correct], y_hat, y_hat_static)
