"""TEST INFRASTRUCTURE ONLY: CPU / fp64 restatement of the Speech2Gesture baseline (reference: scripts/model/speech2gesture.py:9-250,
scripts/train_eval/train_speech2gesture.py:5-37) as plain functions over a state_dict, pinned against the reference modules executed in the
build container by oracle/make_golden_s2g.py (tests/golden/s2g_step.npz; tests/test_oracle_s2g_golden.py).  Only tests/ may import it."""
import torch
import torch.nn.functional as F


def _tf_pad(size, k, s):
    """speech2gesture.py:19-30 (_compute_padding): total SAME padding and whether one more element goes to the far edge"""
    out = (size + s - 1) // s
    total = max(0, (out - 1) * s + k - size)
    return total // 2, total % 2


def conv_tf(x, w, b, stride, padding):
    """Conv2d_tf / Conv1d_tf.forward (speech2gesture.py:32-52, 78-101) on [B,C,H,W] / [B,C,T]"""
    nd = w.dim() - 2
    stride = (stride,) * nd if isinstance(stride, int) else tuple(stride)
    conv = F.conv2d if nd == 2 else F.conv1d
    if padding == 'VALID':
        return conv(x, w, b, stride)
    pads, odd = zip(*[_tf_pad(x.shape[2 + d], w.shape[2 + d], stride[d]) for d in range(nd)])
    if any(odd):
        x = F.pad(x, [0, odd[1], 0, odd[0]] if nd == 2 else [0, odd[0]])
    return conv(x, w, b, stride, padding=pads)


def _bn(x, sd, name, training, updates):
    """BatchNorm{1,2}d; `updates` collects the new running statistics (momentum 0.1, unbiased variance) in train mode"""
    rm, rv = sd[name + '.running_mean'], sd[name + '.running_var']
    if training:
        dims = [0] + list(range(2, x.dim()))
        n = x.numel() // x.shape[1]
        mean = x.mean(dims); var = x.var(dims, unbiased=False)
        if updates is not None:
            updates[name + '.running_mean'] = (0.9 * updates.get(name + '.running_mean', rm) + 0.1 * mean.detach()).to(rm.dtype)
            updates[name + '.running_var'] = (0.9 * updates.get(name + '.running_var', rv) + 0.1 * var.detach() * n / max(n - 1, 1)).to(rv.dtype)
            updates[name + '.num_batches_tracked'] = updates.get(name + '.num_batches_tracked', sd[name + '.num_batches_tracked']) + 1
    else:
        mean, var = rm, rv
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + 1e-5) * sd[name + '.weight'].view(shape) + sd[name + '.bias'].view(shape)


def cnr(x, sd, name, stride, padding, training, updates):
    """ConvNormRelu (speech2gesture.py:104-117): name.0 conv, name.1 norm, LeakyReLU(0.2)"""
    y = conv_tf(x, sd[name + '.0.weight'], sd[name + '.0.bias'], stride, padding)
    return F.leaky_relu(_bn(y, sd, name + '.1', training, updates), 0.2)


def generator_forward(sd, in_spec, pre_poses, n_poses, training, updates=None):
    """Generator.forward + AudioEncoder.forward (speech2gesture.py:160-195, 213-229): [B,128,L], [B,n_pre,D] -> [B,n_poses,D]"""
    a = 'audio_encoder.'
    x = in_spec.unsqueeze(1).to(sd['final_out.weight'].dtype)
    strides = [1, 2, 1, 2, 1, 2, 1, 1]
    for i in range(8):
        x = cnr(x, sd, a + 'first_net.%d' % i, strides[i], 'VALID' if i == 7 else 'SAME', training, updates)
    x = F.interpolate(x, size=(n_poses, 1), mode='bilinear', align_corners=False).squeeze(3)
    x2 = cnr(cnr(x, sd, a + 'down1.0', 1, 'SAME', training, updates), sd, a + 'down1.1', 1, 'SAME', training, updates)
    skips = [x2]
    for i in range(2, 7):
        skips.append(cnr(skips[-1], sd, a + 'down%d' % i, 2, 'SAME', training, updates))
    h = skips[-1]
    for i in range(1, 6):                                      # UnetUp (:120-130)
        s = skips[-1 - i]
        h = torch.repeat_interleave(h, 2, dim=2)[:, :, :s.shape[2]] + s
        h = cnr(h, sd, a + 'up%d.conv' % i, 1, 'SAME', training, updates)
    p = pre_poses.reshape(pre_poses.shape[0], -1).to(h.dtype)
    p = F.linear(p, sd['pre_pose_encoder.0.weight'], sd['pre_pose_encoder.0.bias'])
    p = F.relu(_bn(p, sd, 'pre_pose_encoder.1', training, updates))
    p = F.linear(p, sd['pre_pose_encoder.3.weight'], sd['pre_pose_encoder.3.bias'])
    h = torch.cat((h, p.unsqueeze(2).repeat(1, 1, n_poses)), dim=1)
    for i in range(4):
        h = cnr(h, sd, 'decoder.%d' % i, 1, 'SAME', training, updates)
    return F.conv1d(h, sd['final_out.weight'], sd['final_out.bias']).transpose(1, 2)


def discriminator_forward(sd, x, training, updates=None):
    """Discriminator.forward (speech2gesture.py:244-250): differences its input in time, [B,T,D] -> [B,1,T']"""
    x = (x[:, 1:] - x[:, :-1]).transpose(1, 2)
    h = F.leaky_relu(conv_tf(x, sd['net.0.weight'], sd['net.0.bias'], 2, 'SAME'), 0.2)
    h = cnr(h, sd, 'net.2', 2, 'SAME', training, updates)
    h = cnr(h, sd, 'net.3', 1, 'SAME', training, updates)
    return conv_tf(h, sd['net.4.weight'], sd['net.4.bias'], 1, 'SAME')


def adam(p, g, m, v, step, lr, b1=0.5, b2=0.999, eps=1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
    return p - lr / (1 - b1 ** step) * m / denom, m, v


def train_iter_oracle(gsd, dsd, gopt, dopt, step, in_spec, target, n_pre, w_reg, w_gan, lr_g, lr_d):
    """train_iter_speech2gesture (train_speech2gesture.py:5-37) with Adam (train.py:104-109) as explicit arithmetic.  gsd / dsd: state dicts
    (floating tensors of one dtype); gopt / dopt: {name: (exp_avg, exp_avg_sq)}.  Returns losses, gradients, the updated dicts."""
    gp = {k: v.detach().clone().requires_grad_(True) for k, v in gsd.items() if v.is_floating_point() and 'running' not in k}
    dp = {k: v.detach().clone().requires_grad_(True) for k, v in dsd.items() if v.is_floating_point() and 'running' not in k}
    gfull = dict(gsd); gfull.update(gp)
    dfull = dict(dsd); dfull.update(dp)
    gup, dup = {}, {}
    n_poses = target.shape[1]
    out = generator_forward(gfull, in_spec, target[:, :n_pre], n_poses, True, gup)
    tm = target[:, 1:] - target[:, :-1]
    om = out[:, 1:] - out[:, :-1]
    dis_real = discriminator_forward(dfull, tm, True, dup)
    dis_fake = discriminator_forward(dfull, om.detach(), True, dup)
    dis_error = F.mse_loss(torch.ones_like(dis_real), dis_real) + F.mse_loss(torch.zeros_like(dis_fake), dis_fake)
    dnames = list(dp)
    dgrads = dict(zip(dnames, torch.autograd.grad(dis_error, [dp[k] for k in dnames])))
    new_dsd = dict(dsd); new_dopt = {}
    for k in dnames:
        m, v = dopt.get(k, (torch.zeros_like(dp[k]), torch.zeros_like(dp[k])))
        p, m, v = adam(dp[k].detach(), dgrads[k], m, v, step, lr_d)
        new_dsd[k] = p; new_dopt[k] = (m, v)
    new_dsd.update(dup)
    # generator step against the UPDATED discriminator
    dfull2 = dict(new_dsd)
    dup2 = {}
    l1 = (out - target).abs().mean()
    dis_out = discriminator_forward(dfull2, om, True, dup2)
    gen_error = F.mse_loss(torch.ones_like(dis_out), dis_out)
    loss = w_reg * l1 + w_gan * gen_error
    gnames = list(gp)
    ggrads = dict(zip(gnames, torch.autograd.grad(loss, [gp[k] for k in gnames])))
    new_gsd = dict(gsd); new_gopt = {}
    for k in gnames:
        m, v = gopt.get(k, (torch.zeros_like(gp[k]), torch.zeros_like(gp[k])))
        p, m, v = adam(gp[k].detach(), ggrads[k], m, v, step, lr_g)
        new_gsd[k] = p; new_gopt[k] = (m, v)
    new_gsd.update(gup)
    new_dsd.update(dup2)
    return dict(losses={'loss': w_reg * l1.item(), 'gen': w_gan * gen_error.item(), 'dis': dis_error.item()}, out=out.detach(), g_grads=ggrads,
                d_grads=dgrads, g_sd=new_gsd, d_sd=new_dsd, g_opt=new_gopt, d_opt=new_dopt)
