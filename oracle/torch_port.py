"""CPU timing baseline -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's hot path is pure ``torch.nn`` on the CPU (roko/rnn_model.py:46-59 called from
roko/inference.py:113-117); /root/reference does not exist on the GPU box, so this module
replays the SAME stock-torch operator sequence (embedding gather, permute, two Linear+ReLU,
reshape, the multi-layer bidirectional ``nn.GRU``, Linear, argmax) from a state_dict.  It is what
``bench.py`` times as ``cpu_baseline`` (kind "port") and under ``--impl reference``; it is
checked against the reference-generated golden vectors in tests/test_oracle.py.
Only tests/ and bench.py's CPU legs may import it.
"""
import torch
import torch.nn.functional as F


class TorchCpuPort:
    def __init__(self, state_dict, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in state_dict.items()}
        self.sd = sd
        self.gru = torch.nn.GRU(500, 128, num_layers=3, batch_first=True, bidirectional=True)
        own = self.gru.state_dict()
        for k in own:
            own[k].copy_(sd["gru." + k])
        self.gru.eval()

    @torch.no_grad()
    def forward(self, x_u8):
        sd = self.sd
        x = x_u8.type(torch.LongTensor)                                   # inference.py:113
        h = F.embedding(x, sd["embedding.weight"]).permute((0, 2, 3, 1))  # rnn_model.py:47-48
        h = F.relu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))         # :50
        h = F.relu(F.linear(h, sd["fc2.weight"], sd["fc2.bias"]))         # :53
        h, _ = self.gru(h.reshape(-1, 90, 500))                           # :56-57
        return F.linear(h, sd["fc4.weight"], sd["fc4.bias"])              # :59

    @torch.no_grad()
    def predict(self, x_u8):
        return torch.argmax(self.forward(x_u8), dim=2)                    # inference.py:116
