"""Stock-PyTorch LXMERT feature extractor -- MEASUREMENT HARNESS, NOT PRODUCT.

The LXMERT encoder that surrounds the graph block is outside the hot path (SURVEY.md sections 2 and 8); the
full-iteration benchmark (`bench.py --workload iteration`, BASELINE configs[1] / configs[2]) still needs one so
that "X-GGM train samples/sec" can be measured with the real amount of surrounding compute and the real 0.86 GB
gradient all-reduce.  This file restates the architecture of the reference's `LXRTFeatureExtraction(mode='lxr')`
(src/lxrt/modeling.py:281-620, 888-952, 1078-1093) with plain torch.nn modules and
`F.scaled_dot_product_attention` -- same layer counts (9 language, 5 cross, 5 visual), same widths, same
sub-module attribute names, hence the same state_dict keys and 207.9 M parameters (tests/test_lxmert_standin.py
loads the reference class's own state_dict with strict=True and compares outputs when oracle/_ref is vendored).
No custom kernels: this is the "stock LXMERT" both arms of the iteration benchmark are meant to share.

`visn_fc` (VisualFeatEncoder) is the one piece with a library implementation; pass
`visn_fc_cls=xggm_b200.VisualFeatEncoder` to use it (same parameter names).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Config:
    def __init__(self, vocab_size=30522, hidden_size=768, num_attention_heads=12, intermediate_size=3072,
                 max_position_embeddings=512, type_vocab_size=2, hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, l_layers=9, x_layers=5, r_layers=5, visual_feat_dim=2048,
                 visual_pos_dim=4, initializer_range=0.02):
        self.__dict__.update(locals())
        del self.__dict__["self"]


class Embeddings(nn.Module):                                   # modeling.py:281-313
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size, padding_idx=0)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size, padding_idx=0)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None):
        pos = torch.arange(input_ids.size(1), dtype=torch.long, device=input_ids.device).unsqueeze(0).expand_as(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = self.word_embeddings(input_ids) + self.position_embeddings(pos) + self.token_type_embeddings(token_type_ids)
        return self.dropout(self.LayerNorm(e))


class Attention(nn.Module):                                    # BertAttention, modeling.py:316-374
    def __init__(self, c):
        super().__init__()
        self.heads, self.p = c.num_attention_heads, c.attention_probs_dropout_prob
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(c.hidden_size, c.hidden_size)
        self.value = nn.Linear(c.hidden_size, c.hidden_size)

    def forward(self, x, ctx, mask=None):
        B, T, H = x.shape
        S = ctx.shape[1]
        q = self.query(x).view(B, T, self.heads, -1).transpose(1, 2)
        k = self.key(ctx).view(B, S, self.heads, -1).transpose(1, 2)
        v = self.value(ctx).view(B, S, self.heads, -1).transpose(1, 2)
        m = None if mask is None else mask.to(q.dtype)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=m, dropout_p=self.p if self.training else 0.0)
        return o.transpose(1, 2).reshape(B, T, H)


class AttOutput(nn.Module):                                    # BertAttOutput / BertOutput share this shape
    def __init__(self, c, in_dim=None):
        super().__init__()
        self.dense = nn.Linear(in_dim or c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, h, resid):
        return self.LayerNorm(self.dropout(self.dense(h)) + resid)


class CrossattLayer(nn.Module):                                # modeling.py:391-400
    def __init__(self, c):
        super().__init__()
        self.att = Attention(c)
        self.output = AttOutput(c)

    def forward(self, x, ctx, ctx_mask=None):
        return self.output(self.att(x, ctx, ctx_mask), x)


class SelfattLayer(nn.Module):                                 # modeling.py:403-414
    def __init__(self, c):
        super().__init__()
        self.self = Attention(c)
        self.output = AttOutput(c)

    def forward(self, x, mask):
        return self.output(self.self(x, x, mask), x)


class Intermediate(nn.Module):                                 # modeling.py:417-431 (erf GeLU)
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)

    def forward(self, x):
        return F.gelu(self.dense(x))


class Layer(nn.Module):                                        # BertLayer, modeling.py:448-459
    def __init__(self, c):
        super().__init__()
        self.attention = SelfattLayer(c)
        self.intermediate = Intermediate(c)
        self.output = AttOutput(c, c.intermediate_size)

    def forward(self, x, mask):
        a = self.attention(x, mask)
        return self.output(self.intermediate(a), a)


class XLayer(nn.Module):                                       # LXRTXLayer, modeling.py:469-527
    def __init__(self, c):
        super().__init__()
        self.visual_attention = CrossattLayer(c)
        self.lang_self_att = SelfattLayer(c)
        self.visn_self_att = SelfattLayer(c)
        self.lang_inter = Intermediate(c)
        self.lang_output = AttOutput(c, c.intermediate_size)
        self.visn_inter = Intermediate(c)
        self.visn_output = AttOutput(c, c.intermediate_size)

    def forward(self, lang, lang_mask, visn, visn_mask):
        l_att = self.visual_attention(lang, visn, visn_mask)   # the SAME cross-attention module serves both directions
        v_att = self.visual_attention(visn, lang, lang_mask)
        l_att = self.lang_self_att(l_att, lang_mask)
        v_att = self.visn_self_att(v_att, visn_mask)
        return self.lang_output(self.lang_inter(l_att), l_att), self.visn_output(self.visn_inter(v_att), v_att)


class VisualFeatEncoderTorch(nn.Module):                       # modeling.py:530-556 in stock torch ops
    def __init__(self, config=None, hidden_size=768, hidden_dropout_prob=0.1, feat_dim=2048, pos_dim=4):
        super().__init__()
        if config is not None:
            hidden_size, hidden_dropout_prob = config.hidden_size, config.hidden_dropout_prob
        self.visn_fc = nn.Linear(feat_dim, hidden_size)
        self.visn_layer_norm = nn.LayerNorm(hidden_size, eps=1e-12)
        self.box_fc = nn.Linear(pos_dim, hidden_size)
        self.box_layer_norm = nn.LayerNorm(hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(hidden_dropout_prob)

    def forward(self, visn_input):
        feats, boxes = visn_input
        x = self.visn_layer_norm(self.visn_fc(feats))
        y = self.box_layer_norm(self.box_fc(boxes))
        return self.dropout((x + y) / 2)


class Encoder(nn.Module):                                      # LXRTEncoder, modeling.py:559-611
    def __init__(self, c, visn_fc_cls=None):
        super().__init__()
        cls = visn_fc_cls or VisualFeatEncoderTorch
        self.visn_fc = cls(hidden_size=c.hidden_size, hidden_dropout_prob=c.hidden_dropout_prob,
                           feat_dim=c.visual_feat_dim, pos_dim=c.visual_pos_dim)
        self.layer = nn.ModuleList(Layer(c) for _ in range(c.l_layers))
        self.x_layers = nn.ModuleList(XLayer(c) for _ in range(c.x_layers))
        self.r_layers = nn.ModuleList(Layer(c) for _ in range(c.r_layers))

    def forward(self, lang, lang_mask, visn_feats, visn_mask=None):
        visn = self.visn_fc(visn_feats)
        if visn.dtype != lang.dtype:        # a library visn_fc returns fp32 whatever the autocast state
            visn = visn.to(lang.dtype)
        for m in self.layer:
            lang = m(lang, lang_mask)
        for m in self.r_layers:
            visn = m(visn, visn_mask)
        for m in self.x_layers:
            lang, visn = m(lang, lang_mask, visn, visn_mask)
        return lang, visn


class Pooler(nn.Module):                                       # modeling.py:614-627
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)

    def forward(self, h):
        return torch.tanh(self.dense(h[:, 0]))


class LXRTModelTorch(nn.Module):                               # LXRTModel, modeling.py:888-952
    def __init__(self, c, visn_fc_cls=None):
        super().__init__()
        self.embeddings = Embeddings(c)
        self.encoder = Encoder(c, visn_fc_cls)
        self.pooler = Pooler(c)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, visual_feats=None):
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        ext = (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0     # modeling.py:919-927
        emb = self.embeddings(input_ids, token_type_ids)
        lang, visn = self.encoder(emb, ext.to(emb.dtype), visual_feats)
        return (lang, visn), self.pooler(lang)


class LXRTFeatureExtractionTorch(nn.Module):
    """state_dict keys = the reference's `LXRTFeatureExtraction` (`bert.*`).  forward -> ((lang, visn), pooled)."""

    def __init__(self, config=None, visn_fc_cls=None):
        super().__init__()
        c = config or Config()
        self.config = c
        self.bert = LXRTModelTorch(c, visn_fc_cls)
        self.apply(self._init)

    def _init(self, m):                                        # init_bert_weights, modeling.py:734-747
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, visual_feats=None):
        return self.bert(input_ids, token_type_ids, attention_mask, visual_feats)


def synthetic_language(B, T=20, vocab=30522, seed=0):
    """[CLS] + L random word pieces + [SEP] + zero padding, L ~ U{3..18}; mask accordingly."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(B, T, dtype=torch.long)
    mask = torch.zeros(B, T, dtype=torch.long)
    L = torch.randint(3, T - 1, (B,), generator=g)
    for b in range(B):
        n = int(L[b])
        ids[b, 0] = 101
        ids[b, 1:1 + n] = torch.randint(1000, 30000, (n,), generator=g)
        ids[b, 1 + n] = 102
        mask[b, :n + 2] = 1
    return ids, mask


if __name__ == "__main__":
    m = LXRTFeatureExtractionTorch()
    print(sum(p.numel() for p in m.parameters()))
