"""NGCF -- drop-in for the reference's recommender/NGCF.py (train loop :31-79,
encoder :163-212).  Propagation runs on the agcf SpMM kernel (forward and backward),
the d x d weight products are library GEMMs; the loop itself is the reference's."""
import torch

from ..encoder import NGCF_Encoder, TorchGraphInterface  # noqa: F401
from ..util.loss import bpr_l2_fused, bpr_loss, l2_reg_loss  # noqa: F401
from ._base import GraphRecommender


class NGCF(GraphRecommender):
    model_name = "NGCF"

    def _build_model(self):
        return NGCF_Encoder(self.data, self.args.emb_size, self.args.n_layers)

    def train(self, requires_adjgrad=False, requires_embgrad=False, gradIterationNum=10, Epoch=0, optimizer=None,
              evalNum=5):
        self.bestPerformance = []
        model = self.model.cuda()
        if optimizer is None:
            optimizer = torch.optim.Adam(model.parameters(), lr=self.args.lRate)
        self._grad_buffers(requires_adjgrad, requires_embgrad, model)
        maxEpoch = Epoch if Epoch else self.args.maxEpoch
        dev = model.embedding_dict['user_emb'].device
        for epoch in range(maxEpoch):
            for n, batch in enumerate(self._epoch_batches(dev)):
                user_idx, pos_idx, neg_idx = batch
                model.train()
                rec_user_emb, rec_item_emb = model()
                # the three gathers + bpr_loss + l2_reg_loss(reg, user_emb, pos_item_emb) of the reference, one fused op
                batch_loss = bpr_l2_fused(rec_user_emb, rec_item_emb, user_idx, pos_idx, neg_idx, self.args.reg)
                optimizer.zero_grad()
                batch_loss.backward()
                self._accumulate_grads(requires_adjgrad, requires_embgrad, maxEpoch, epoch, gradIterationNum)
                optimizer.step()
                if n % 100 == 0:
                    print('training:', epoch + 1, 'batch', n, 'batch_loss:', batch_loss.item())
            model.eval()
            with torch.no_grad():
                self.user_emb, self.item_emb = self.model()
            if epoch % evalNum == 0:
                self.evaluate(epoch)
        self.user_emb, self.item_emb = self.best_user_emb, self.best_item_emb
        return self._train_returns(requires_adjgrad, requires_embgrad)
