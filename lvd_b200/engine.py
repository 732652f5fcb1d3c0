"""UNet3D denoiser engine: the layer graph of the reference's UNet3DConditionModel.forward
(models/unet_3d_condition.py:642-859) sequenced over the C-ABI kernels.

Host code here is plumbing only -- buffer allocation, weight repacking at load, launch order.  Activations are
channels-last fp16 row matrices [(b*F+f)*h*w + p, C]; in that layout the reference's permutes
(transformer_2d.py:318-327, transformer_temporal.py:142-156,176-182) are identities on memory, skip concatenation is
a column slice, and every Linear / Conv2d / Conv3d is one `lvd_gemm` launch.
"""
import torch

from . import ops, repack


class LayerWeights(dict):
    __getattr__ = dict.__getitem__


def _h(t, dev):
    return t.detach().to(device=dev, dtype=torch.float16).contiguous()


class UNetEngine:
    """Holds repacked fp16 weights on one GPU and runs the forward.  `cfg` mirrors the reference's register_to_config
    fields (block_out_channels, layers_per_block, cross_attention_dim, attention_head_dim, norm_num_groups, norm_eps,
    gated)."""

    def __init__(self, cfg, state_dict, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("lvd_b200.UNetEngine needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = dict(cfg)
        self.dev = torch.device(device)
        self.w = {}
        self.fuser_enabled = True
        self._load(state_dict)

    # ------------------------------------------------------------------------------------------------ weights
    def _load(self, sd):
        dev = self.dev
        w = self.w
        g = lambda k: _h(sd[k], dev)  # noqa: E731
        cfg = self.cfg

        def lin(p, bias=True):
            return LayerWeights(w=g(p + ".weight"), b=g(p + ".bias") if bias and (p + ".bias") in sd else None)

        def norm(p):
            return LayerWeights(w=g(p + ".weight"), b=g(p + ".bias"))

        def conv3(p):
            return LayerWeights(w=repack.conv_weight_krsc(g(p + ".weight")), b=g(p + ".bias"))

        def block(p, cross):
            d = LayerWeights(norm1=norm(p + ".norm1"), norm2=norm(p + ".norm2"), norm3=norm(p + ".norm3"))
            d["qkv1"] = torch.cat([g(p + ".attn1.to_q.weight"), g(p + ".attn1.to_k.weight"),
                                   g(p + ".attn1.to_v.weight")], 0).contiguous()
            d["out1"] = lin(p + ".attn1.to_out.0")
            if cross:
                d["q2"] = g(p + ".attn2.to_q.weight")
                d["kv2"] = torch.cat([g(p + ".attn2.to_k.weight"), g(p + ".attn2.to_v.weight")], 0).contiguous()
            else:
                d["qkv2"] = torch.cat([g(p + ".attn2.to_q.weight"), g(p + ".attn2.to_k.weight"),
                                       g(p + ".attn2.to_v.weight")], 0).contiguous()
            d["out2"] = lin(p + ".attn2.to_out.0")
            wi, bi = repack.geglu_interleave(g(p + ".ff.net.0.proj.weight"), g(p + ".ff.net.0.proj.bias"))
            d["ff1"] = LayerWeights(w=wi, b=bi)
            d["ff2"] = lin(p + ".ff.net.2")
            # LayerNorm folded into the linear that consumes it (ops.ln_fold; models/attention.py:126-170)
            d["qkv1_ln"] = ops.ln_fold(d["qkv1"], None, d["norm1"].w, d["norm1"].b)
            k2 = "q2" if cross else "qkv2"
            d[k2 + "_ln"] = ops.ln_fold(d[k2], None, d["norm2"].w, d["norm2"].b)
            d["ff1_ln"] = ops.ln_fold(wi, bi, d["norm3"].w, d["norm3"].b)
            f = p + ".fuser"
            if (f + ".linear.weight") in sd:  # GatedSelfAttentionDense (models/attention.py:26-60)
                fwi, fbi = repack.geglu_interleave(g(f + ".ff.net.0.proj.weight"), g(f + ".ff.net.0.proj.bias"))
                d["fuser"] = LayerWeights(
                    linear=lin(f + ".linear"), norm1=norm(f + ".norm1"), norm2=norm(f + ".norm2"),
                    qkv=torch.cat([g(f + ".attn.to_q.weight"), g(f + ".attn.to_k.weight"),
                                   g(f + ".attn.to_v.weight")], 0).contiguous(),
                    out=lin(f + ".attn.to_out.0"), ff1=LayerWeights(w=fwi, b=fbi), ff2=lin(f + ".ff.net.2"),
                    g_attn=float(torch.tanh(sd[f + ".alpha_attn"].float())),
                    g_dense=float(torch.tanh(sd[f + ".alpha_dense"].float())))
            return d

        def transformer(p, cross):
            return LayerWeights(norm=norm(p + ".norm"), proj_in=lin(p + ".proj_in"),
                                block=block(p + ".transformer_blocks.0", cross), proj_out=lin(p + ".proj_out"))

        self._temb_slices = {}
        temb_w, temb_b = [], []
        off = 0

        def resnet(p):
            nonlocal off
            d = LayerWeights(norm1=norm(p + ".norm1"), conv1=conv3(p + ".conv1"), norm2=norm(p + ".norm2"),
                             conv2=conv3(p + ".conv2"))
            tw = g(p + ".time_emb_proj.weight")
            temb_w.append(tw)
            temb_b.append(g(p + ".time_emb_proj.bias"))
            d["temb"] = (off, off + tw.shape[0])
            off += tw.shape[0]
            if (p + ".conv_shortcut.weight") in sd:
                sw = g(p + ".conv_shortcut.weight")
                d["shortcut"] = LayerWeights(w=sw.reshape(sw.shape[0], sw.shape[1]).contiguous(),
                                             b=g(p + ".conv_shortcut.bias"))
            return d

        def tconv(p):
            out = []
            for name, idx in (("conv1", 2), ("conv2", 3), ("conv3", 3), ("conv4", 3)):
                out.append(LayerWeights(norm=norm(f"{p}.{name}.0"), w=repack.tconv_weight(g(f"{p}.{name}.{idx}.weight")),
                                        b=g(f"{p}.{name}.{idx}.bias")))
            return out

        boc = list(cfg["block_out_channels"])
        nb = len(boc)
        w["conv_in"] = LayerWeights(w=g("conv_in.weight"), b=g("conv_in.bias"),
                                    w_pad=repack.conv_in_weight_padded(g("conv_in.weight")))
        w["time1"] = lin("time_embedding.linear_1")
        w["time2"] = lin("time_embedding.linear_2")
        w["transformer_in"] = transformer("transformer_in", cross=False)
        w["down"] = []
        for i in range(nb):
            p = f"down_blocks.{i}"
            layers = []
            for j in range(cfg["layers_per_block"]):
                L = LayerWeights(resnet=resnet(f"{p}.resnets.{j}"), tconv=tconv(f"{p}.temp_convs.{j}"))
                if i < nb - 1:
                    L["attn"] = transformer(f"{p}.attentions.{j}", cross=True)
                    L["tattn"] = transformer(f"{p}.temp_attentions.{j}", cross=False)
                layers.append(L)
            blk = LayerWeights(layers=layers)
            if i < nb - 1:
                blk["down"] = conv3(f"{p}.downsamplers.0.conv")
            w["down"].append(blk)
        w["mid"] = LayerWeights(
            res0=resnet("mid_block.resnets.0"), tconv0=tconv("mid_block.temp_convs.0"),
            attn=transformer("mid_block.attentions.0", cross=True),
            tattn=transformer("mid_block.temp_attentions.0", cross=False),
            res1=resnet("mid_block.resnets.1"), tconv1=tconv("mid_block.temp_convs.1"))
        w["up"] = []
        for i in range(nb):
            p = f"up_blocks.{i}"
            layers = []
            for j in range(cfg["layers_per_block"] + 1):
                L = LayerWeights(resnet=resnet(f"{p}.resnets.{j}"), tconv=tconv(f"{p}.temp_convs.{j}"))
                if i > 0:
                    L["attn"] = transformer(f"{p}.attentions.{j}", cross=True)
                    L["tattn"] = transformer(f"{p}.temp_attentions.{j}", cross=False)
                layers.append(L)
            blk = LayerWeights(layers=layers)
            if i < nb - 1:
                blk["up"] = conv3(f"{p}.upsamplers.0.conv")
            w["up"].append(blk)
        if "position_net.linears.0.weight" in sd:
            w["position_net"] = LayerWeights(l0=lin("position_net.linears.0"), l2=lin("position_net.linears.2"),
                                             l4=lin("position_net.linears.4"),
                                             null_pos=g("position_net.null_positive_feature"),
                                             null_xyxy=g("position_net.null_position_feature"))
        w["norm_out"] = norm("conv_norm_out")
        wko = repack.conv_weight_krsc(g("conv_out.weight"))
        wpad, bpad = repack.pad_out_channels(wko, g("conv_out.bias"))  # conv_out runs as one 64-wide GEMM tile
        w["conv_out"] = LayerWeights(w=wko, b=g("conv_out.bias"), w_pad=wpad, b_pad=bpad, cout=wko.shape[0])
        # all 22 time_emb_proj Linear layers fused into one GEMM per forward (ResnetBlock2D)
        tw = torch.cat(temb_w, 0)
        pad = (-tw.shape[0]) % 64
        if pad:
            tw = torch.cat([tw, torch.zeros(pad, tw.shape[1], dtype=tw.dtype, device=dev)], 0)
        tb = torch.cat(temb_b + ([torch.zeros(pad, dtype=tw.dtype, device=dev)] if pad else []), 0)
        w["temb_all"] = LayerWeights(w=tw.contiguous(), b=tb.contiguous())

    # ------------------------------------------------------------------------------------------------ blocks
    def _resnet(self, L, x, n_img, h, w_, temb_all, out=None, gn_next=None):
        """`gn_next` = samples of the GroupNorm that reads this block's output NEXT (the 5-D norm of the temporal conv
        layer): conv2's epilogue then produces its statistics; conv1's epilogue produces norm2's."""
        eps = self.cfg["norm_eps"]
        t = ops.groupnorm(x, n_img, L.norm1.w, L.norm1.b, eps, silu=True)
        lo, hi = L.temb
        t = ops.conv2d_3x3(t, n_img, h, w_, L.conv1.w, L.conv1.b, rowvec=temb_all[:, lo:hi], gn=(n_img, 32))
        t = ops.groupnorm(t, n_img, L.norm2.w, L.norm2.b, eps, silu=True)
        sc = x
        if "shortcut" in L:
            sc = ops.linear(x, L.shortcut.w, L.shortcut.b)
        return ops.conv2d_3x3(t, n_img, h, w_, L.conv2.w, L.conv2.b, residual=sc, out=out,
                              gn=(gn_next, 32) if gn_next else None)

    def _tconv(self, layers, x, batch, frames, hw):
        t = x
        for i, L in enumerate(layers):
            t = ops.groupnorm(t, batch, L.norm.w, L.norm.b, 1e-5, silu=True)
            t = ops.conv_temporal(t, batch, frames, hw, L.w, L.b, residual=x if i == 3 else None,
                                  gn=(batch, 32) if i < 3 else None)  # the next norm of the chain reads this output
        return t

    def _fuser(self, Fz, x, objs, *, heads, n_img, hw):
        """GatedSelfAttentionDense.forward (models/attention.py:44-60): x += tanh(a_attn) * SA(LN([x; W objs]))[:hw];
        x += tanh(a_dense) * FF(LN(x))."""
        c = x.shape[1]
        n_obj = objs.shape[0] // n_img
        seq = hw + n_obj
        lo = ops.linear(objs, Fz.linear.w, Fz.linear.b)
        cat = torch.empty((n_img * seq, c), dtype=torch.float16, device=self.dev)
        ops.copy_rows_seq(x, cat, hw, src_seq=hw, src_off=0, dst_seq=seq, dst_off=0, groups=n_img)
        ops.copy_rows_seq(lo, cat, n_obj, src_seq=n_obj, src_off=0, dst_seq=seq, dst_off=hw, groups=n_img)
        n1 = ops.layernorm(cat, Fz.norm1.w, Fz.norm1.b)
        qkv = ops.linear(n1, Fz.qkv)
        a = ops.attention(qkv, qkv, heads=heads, nz=n_img, s_q=seq, s_k=seq, k_col0=c, v_col0=2 * c)
        av = torch.empty((n_img * hw, c), dtype=torch.float16, device=self.dev)
        ops.copy_rows_seq(a, av, hw, src_seq=seq, src_off=0, dst_seq=hw, dst_off=0, groups=n_img)
        x = ops.linear(av, Fz.out.w, Fz.out.b, gamma=Fz.g_attn, residual=x)
        n2 = ops.layernorm(x, Fz.norm2.w, Fz.norm2.b)
        gg = ops.linear(n2, Fz.ff1.w, Fz.ff1.b, act=ops.ACT_GEGLU)
        return ops.linear(gg, Fz.ff2.w, Fz.ff2.b, gamma=Fz.g_dense, residual=x)

    def _position_net(self, gligen):
        """PositionNet.forward, text-only features (models/unet_3d_condition.py:119-150)."""
        P = self.w["position_net"]
        h = ops.gligen_prep(gligen["boxes"], gligen["masks"], gligen["positive_embeddings"], P.null_pos, P.null_xyxy)
        h = ops.linear(h, P.l0.w, P.l0.b, act=ops.ACT_SILU)
        h = ops.linear(h, P.l2.w, P.l2.b, act=ops.ACT_SILU)
        return ops.linear(h, P.l4.w, P.l4.b)

    @staticmethod
    def _ln_linear(t, norm, w, b, folded, **kw):
        """linear(LayerNorm(t)): on the raw t with the LayerNorm folded into the weight when t's producer left its row
        statistics, through the LayerNorm kernel otherwise."""
        st = ops.ln_rowstats(t)
        if st is not None:
            return ops.linear(t, folded.w, ln=(st, folded), **kw)
        return ops.linear(ops.layernorm(t, norm.w, norm.b), w, b, **kw)

    def _block(self, B, t, *, heads, spatial, n_img, hw, batch, frames, ctx_kv=None, key=None, capture=None,
               objs=None):
        c = t.shape[1]
        packed = (not spatial) and frames % 8 == 0 and frames <= 32
        fm2sm = ("fm2sm", batch, frames, hw) if packed else None
        sm2fm = ("sm2fm", batch, frames, hw) if packed else None
        qkv = self._ln_linear(t, B.norm1, B.qkv1, None, B.qkv1_ln, row_map=fm2sm)
        if spatial:
            a = ops.attention(qkv, qkv, heads=heads, nz=n_img, s_q=hw, s_k=hw, k_col0=c, v_col0=2 * c)
        else:
            a = ops.attention(qkv, qkv, heads=heads, nz=batch, temporal=True, frames=frames, hw=hw, k_col0=c,
                              v_col0=2 * c, packed=packed)
        fused = spatial and objs is not None and "fuser" in B and self.fuser_enabled
        t = ops.linear(a, B.out1.w, B.out1.b, residual=t, row_map=sm2fm, ln_emit=not fused)
        if fused:
            t = self._fuser(B.fuser, t, objs, heads=heads, n_img=n_img, hw=hw)
        if spatial:
            q = self._ln_linear(t, B.norm2, B.q2, None, B.q2_ln)
            kv = ops.linear(ctx_kv, B.kv2)
            extra = {}
            if capture is not None:
                extra = capture(key, n_img, heads, hw)
            a = ops.attention(q, kv, heads=heads, nz=n_img, s_q=hw, s_k=ctx_kv.shape[0] // batch, k_col0=0, v_col0=c,
                              kv_z_div=frames, **extra)
        else:
            qkv = self._ln_linear(t, B.norm2, B.qkv2, None, B.qkv2_ln, row_map=fm2sm)
            a = ops.attention(qkv, qkv, heads=heads, nz=batch, temporal=True, frames=frames, hw=hw, k_col0=c,
                              v_col0=2 * c, packed=packed)
        t = ops.linear(a, B.out2.w, B.out2.b, residual=t, row_map=sm2fm if not spatial else None, ln_emit=True)
        gg = self._ln_linear(t, B.norm3, B.ff1.w, B.ff1.b, B.ff1_ln, act=ops.ACT_GEGLU)
        return ops.linear(gg, B.ff2.w, B.ff2.b, residual=t)

    def _transformer(self, T, x, *, spatial, n_img, hw, batch, frames, ctx_kv=None, key=None, capture=None,
                     out=None, objs=None):
        samples = n_img if spatial else batch
        t = ops.groupnorm(x, samples, T.norm.w, T.norm.b, 1e-6, silu=False)
        t = ops.linear(t, T.proj_in.w, T.proj_in.b, ln_emit=True)
        heads = t.shape[1] // self.cfg["attention_head_dim"]
        t = self._block(T.block, t, heads=heads, spatial=spatial, n_img=n_img, hw=hw, batch=batch, frames=frames,
                        ctx_kv=ctx_kv, key=key, capture=capture, objs=objs)
        return ops.linear(t, T.proj_out.w, T.proj_out.b, residual=x, out=out)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, sample, timestep, encoder_hidden_states, *, capture=None, stop_after_key=None, gligen=None):
        """sample (B, 4, F, h, w) fp16 cuda; timestep scalar / 0-d / (B,) ; encoder_hidden_states (B, 77, D) fp16.
        `capture(key, n_img, heads, hw) -> dict(p_full=..., p_tok=..., tok_cols=...)` is consulted for every
        cross-attention key (kind, stage, layer, 0)."""
        cfg = self.cfg
        w = self.w
        dev = self.dev
        ops.gn_reset(dev)
        boc = list(cfg["block_out_channels"])
        nb = len(boc)
        sample = sample.to(device=dev, dtype=torch.float16).contiguous()
        b, cin, frames, h, w_ = sample.shape
        n_img = b * frames
        ehs = encoder_hidden_states.to(device=dev, dtype=torch.float16).contiguous()
        ctx = ehs.reshape(b * ehs.shape[1], ehs.shape[2])
        if torch.is_tensor(timestep) and timestep.is_cuda and timestep.dtype == torch.float32 and \
                timestep.numel() == n_img:
            t = timestep  # pre-expanded device tensor (CUDA-graph replay: the value is read at run time)
        else:
            t = torch.as_tensor(timestep, device=dev).reshape(-1).to(torch.float32)
            t = t.expand(b).repeat_interleave(frames).contiguous()
        temb = ops.timestep_embed(t, boc[0])
        temb = ops.linear(temb, w["time1"].w, w["time1"].b, act=ops.ACT_SILU)  # silu fused: only silu(temb@W1) is used
        temb = ops.linear(temb, w["time2"].w, w["time2"].b)
        temb_all = ops.linear(ops.silu(temb), w["temb_all"].w, w["temb_all"].b)

        x = ops.conv_in_tc(sample, w["conv_in"].w_pad, w["conv_in"].b)
        hw = h * w_
        x = self._transformer(w["transformer_in"], x, spatial=False, n_img=n_img, hw=hw, batch=b, frames=frames)
        objs = None
        if gligen is not None:
            if "position_net" not in w:
                raise ValueError("`gligen` inputs given but the model has no GLIGEN weights (attention_type='default')")
            objs = self._position_net(gligen)  # unet_3d_condition.py:739-748

        class _Stop(Exception):
            pass

        def layer(L, x, kind, stage, j, h, w_):
            hw = h * w_
            x = self._resnet(L.resnet, x, n_img, h, w_, temb_all, gn_next=b)
            x = self._tconv(L.tconv, x, b, frames, hw)
            if "attn" in L:
                key = (kind, stage, j, 0)
                x = self._transformer(L.attn, x, spatial=True, n_img=n_img, hw=hw, batch=b, frames=frames, ctx_kv=ctx,
                                      key=key, capture=capture, objs=objs)
                if stop_after_key is not None and key == tuple(stop_after_key):
                    raise _Stop()
                x = self._transformer(L.tattn, x, spatial=False, n_img=n_img, hw=hw, batch=b, frames=frames)
            return x

        try:
            skips = [(x, h, w_)]
            for i in range(nb):
                blk = w["down"][i]
                for j, L in enumerate(blk.layers):
                    x = layer(L, x, "down", i, j, h, w_)
                    skips.append((x, h, w_))
                if "down" in blk:
                    x = ops.conv2d_3x3_s2(x, n_img, h, w_, blk.down.w, blk.down.b)
                    h, w_ = h // 2, w_ // 2
                    skips.append((x, h, w_))
            M = w["mid"]
            hw = h * w_
            x = self._resnet(M.res0, x, n_img, h, w_, temb_all, gn_next=b)
            x = self._tconv(M.tconv0, x, b, frames, hw)
            x = self._transformer(M.attn, x, spatial=True, n_img=n_img, hw=hw, batch=b, frames=frames, ctx_kv=ctx,
                                  key=("mid", 0, 0, 0), capture=capture, objs=objs)
            x = self._transformer(M.tattn, x, spatial=False, n_img=n_img, hw=hw, batch=b, frames=frames)
            x = self._resnet(M.res1, x, n_img, h, w_, temb_all, gn_next=b)
            x = self._tconv(M.tconv1, x, b, frames, hw)
            for i in range(nb):
                blk = w["up"][i]
                for j, L in enumerate(blk.layers):
                    skip, sh, sw = skips.pop()
                    assert (sh, sw) == (h, w_)
                    cat = torch.empty((x.shape[0], x.shape[1] + skip.shape[1]), dtype=torch.float16, device=dev)
                    ops.copy_rows(x, cat[:, :x.shape[1]])
                    ops.copy_rows(skip, cat[:, x.shape[1]:])
                    x = layer(L, cat, "up", i, j, h, w_)
                if "up" in blk:
                    x = ops.upsample2x(x, n_img, h, w_)
                    h, w_ = 2 * h, 2 * w_
                    x = ops.conv2d_3x3(x, n_img, h, w_, blk.up.w, blk.up.b)
        except _Stop:
            return None
        x = ops.groupnorm(x, n_img, w["norm_out"].w, w["norm_out"].b, cfg["norm_eps"], silu=True)
        co = w["conv_out"]
        return ops.conv_out_tc(x, co.w_pad, co.b_pad, co.cout, b, frames, h, w_)
