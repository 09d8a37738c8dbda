"""The reference learner's math in plain PyTorch fp32 on the CPU — TEST INFRASTRUCTURE (the fp32 reference the
floating-point training kernel is compared with, and bench_train.py's CPU leg).  The reference runs these very libtorch
ops through tch (synthesis/src/alpha_zero.rs:73-92; network study-connect4/src/policies.rs:13-46; Adam = tch
nn::Adam::default() with set_weight_decay, alpha_zero.rs:33-36)."""
import numpy as np
import torch

DIMS = (63, 128, 96, 64, 48, 12)


class TorchLearner:
    def __init__(self, blob, lr, weight_decay=0.0, policy_weight=1.0, value_weight=1.0, batch_size=32):
        torch.set_num_threads(1)  # study-connect4/src/main.rs:85-86
        blob = np.asarray(blob, np.float32)
        self.params, off = [], 0
        for i, o in zip(DIMS[:-1], DIMS[1:]):
            w = torch.tensor(blob[off:off + i * o].reshape(o, i).copy(), requires_grad=True); off += i * o
            b = torch.tensor(blob[off:off + o].copy(), requires_grad=True); off += o
            self.params += [w, b]
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
        self.pw, self.vw, self.batch_mean = policy_weight, value_weight, 1.0 / batch_size

    def set_lr(self, lr):
        for g in self.opt.param_groups:
            g["lr"] = lr

    def forward(self, x):
        for l in range(5):
            x = torch.nn.functional.linear(x, self.params[2 * l], self.params[2 * l + 1])
            if l < 4:
                x = torch.relu(x)
        return x[:, :9], x[:, 9:]

    def step(self, states, target_pi, target_v):
        pi_logits, v_logits = self.forward(states.reshape(states.shape[0], -1))
        log_pi = torch.log_softmax(pi_logits, -1)
        log_v = torch.log_softmax(v_logits, -1)
        pi_loss = self.batch_mean * torch.nn.functional.kl_div(log_pi, target_pi, reduction="sum", log_target=False)
        v_loss = self.batch_mean * torch.nn.functional.kl_div(log_v, target_v, reduction="sum", log_target=False)
        loss = self.pw * pi_loss + self.vw * v_loss
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(pi_loss), float(v_loss)

    def run(self, states, pis, vs, batch_index):
        st, pi, v = torch.from_numpy(np.ascontiguousarray(states)), torch.from_numpy(np.ascontiguousarray(pis)), torch.from_numpy(np.ascontiguousarray(vs))
        out = np.zeros((len(batch_index), 2), np.float32)
        for k, idx in enumerate(batch_index):
            ii = torch.from_numpy(np.asarray(idx, np.int64))
            out[k] = self.step(st.index_select(0, ii), pi.index_select(0, ii), v.index_select(0, ii))
        return out

    def blob(self):
        return np.concatenate([p.detach().numpy().reshape(-1) for p in self.params]).astype(np.float32)
