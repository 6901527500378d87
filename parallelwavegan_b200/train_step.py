"""One HiFi-GAN-style GAN train step (bin/train.py:189-340 `Trainer._train_step`), re-hosted on the
libpwgb forward / backward kernels.  Data parallelism = one process per GPU with
``torch.nn.parallel.DistributedDataParallel`` (bucketed NCCL gradient all-reduce overlapped with the
backward), replacing the reference's apex DDP (train.py:1494-1503)."""
import torch


class GanTrainStep:
    def __init__(self, generator, discriminator, criterion, opt_g, opt_d, lambda_aux=45.0, lambda_adv=1.0,
                 lambda_feat_match=2.0, grad_norm_g=-1, grad_norm_d=-1):
        self.g, self.d = generator, discriminator
        self.crit = criterion  # dict: mel, gen_adv, dis_adv, feat_match (any subset like the reference configs)
        self.opt_g, self.opt_d = opt_g, opt_d
        self.lambda_aux, self.lambda_adv, self.lambda_fm = lambda_aux, lambda_adv, lambda_feat_match
        self.grad_norm_g, self.grad_norm_d = grad_norm_g, grad_norm_d

    @staticmethod
    def _params(m):
        return (m.module if hasattr(m, "module") else m).parameters()

    def __call__(self, c, y):
        """c: (B, mels, frames), y: (B, 1, T).  Returns a dict of loss tensors (device scalars, no host sync)."""
        stats = {}
        # ---------------- generator phase (train.py:200-295)
        y_ = self.g(c)
        gen_loss = 0.0
        if "mel" in self.crit:
            mel = self.crit["mel"](y_, y)
            stats["mel_loss"] = mel.detach()
            gen_loss = gen_loss + self.lambda_aux * mel
        # D's weight gradients of this phase are discarded by the reference (optD.zero_grad at train.py:327): skip them
        for p in self._params(self.d):
            p.requires_grad_(False)
        p_ = self.d(y_)
        adv = self.crit["gen_adv"](p_)
        stats["adversarial_loss"] = adv.detach()
        gen_loss = gen_loss + self.lambda_adv * adv
        if "feat_match" in self.crit:
            with torch.no_grad():
                p = self.d(y)
            fm = self.crit["feat_match"](p_, p)
            stats["feature_matching_loss"] = fm.detach()
            gen_loss = gen_loss + self.lambda_fm * fm
        self.opt_g.zero_grad(set_to_none=True)
        gen_loss.backward()
        if self.grad_norm_g > 0:
            torch.nn.utils.clip_grad_norm_(self._params(self.g), self.grad_norm_g)
        self.opt_g.step()
        for p in self._params(self.d):
            p.requires_grad_(True)
        # ---------------- discriminator phase (train.py:300-335), with the updated generator
        with torch.no_grad():
            y_ = self.g(c)
        p = self.d(y)
        p_ = self.d(y_.detach())
        real, fake = self.crit["dis_adv"](p_, p)
        dis_loss = real + fake
        stats["real_loss"], stats["fake_loss"] = real.detach(), fake.detach()
        self.opt_d.zero_grad(set_to_none=True)
        dis_loss.backward()
        if self.grad_norm_d > 0:
            torch.nn.utils.clip_grad_norm_(self._params(self.d), self.grad_norm_d)
        self.opt_d.step()
        stats["generator_loss"] = gen_loss.detach()
        stats["discriminator_loss"] = dis_loss.detach()
        return stats
