"""Personalisation loop on the B200 training step: the host loop of
iPERCore/services/personalization.py:95-151 (``Personalizer.run``) over ``ipercore_b200.train.LWGTrainStep``.

Same schedule — ``(niters_or_epochs_no_decay + niters_or_epochs_decay) * num_videos`` iterations (deploy.toml:99-103), G updated on
every ``train_G_every_n_iterations``-th batch and D on every batch (lwg_trainer.py:326-352), linear learning-rate decay over the
decay phase (lwg_trainer.py:300-324) — and the same product: ``personalized.pth`` = the generator's ``state_dict`` under the
reference's names, which both the reference's ``Imitator`` and ``ipercore_b200.generator.AttentionLWBGenerator`` load.

What stays upstream: the dataset (``ProcessedVideoDataset``) and ``LWGTrainer.set_input`` — the batches handed to ``run`` are the
tensors ``set_input`` produces (bg_inputs, src_inputs, tsf_inputs, Tst, real_src, real_tsf, real_bg, body_mask)."""
import os

import torch


def run(step, batches, num_videos=1, niters_no_decay=100, niters_decay=0, train_G_every_n_iterations=1, lr=1e-4, final_lr=1e-6,
        ckpt_path=None, on_iter=None):
    """step: LWGTrainStep; batches: an iterable of batch dicts that is cycled (the reference cycles its DataLoader per epoch).
    Returns the list of per-iteration loss dicts (device tensors)."""
    total_iters = (niters_no_decay + niters_decay) * num_videos
    history, total, cur_lr = [], 0, lr
    step.set_lr(cur_lr)
    while total < total_iters:
        seen = 0
        for i_batch, batch in enumerate(batches):
            trainable = (i_batch + 1) % train_G_every_n_iterations == 0
            out = step.step(batch, trainable=trainable)
            history.append(out)
            total += 1; seen += 1
            if on_iter is not None:
                on_iter(total, out)
            if niters_decay > 0 and total > niters_no_decay * num_videos:        # lwg_trainer.py:300-324, once per decay iteration
                cur_lr = max(final_lr, cur_lr - (lr - final_lr) / (niters_decay * num_videos))
                step.set_lr(cur_lr)
            if total >= total_iters:
                break
        if seen == 0:
            raise ValueError("personalize.run: the batch iterable is empty")
    if ckpt_path is not None:
        os.makedirs(os.path.dirname(os.path.abspath(ckpt_path)), exist_ok=True)
        torch.save({k: v.detach().float().cpu().clone() for k, v in step.G.net.state_dict().items()}, ckpt_path)
    return history
