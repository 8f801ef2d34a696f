import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import ref_loader, port_modules as P
U, Pol = ref_loader.load_ref_models()
torch.manual_seed(0)
B, L, C = 5, 12, 6
def close(a, b, name, tol=1e-5):
    d = (a-b).abs().max().item(); print(f"{name}: {d:.3e}"); assert d < tol, name
# encoder variants
for (E,H,bi,nl) in [(256,512,True,1),(300,256,True,2),(256,512,False,1)]:
    enc = U.EncoderLSTM(992, E, H, 0, 0.5, bi, nl).eval()
    toks = torch.randint(4, 992, (B, L)); lens = torch.tensor([12, 10, 7, 7, 3])
    for i,l in enumerate(lens): toks[i, l:] = 0
    ctx, h, c = enc(toks, lens)
    sd = enc.state_dict()
    ctx2, h2, c2 = P.encoder_lstm(sd, toks, lens, bidirectional=bi, num_layers=nl, drop_ratio=0.5)
    close(ctx, ctx2, "enc ctx"); close(h, h2, "enc h"); close(c, c2, "enc c")
    # train mode w/ torch RNG
    if nl == 1:
        enc.train(); torch.manual_seed(5); ctx, h, c = enc(toks, lens)
        torch.manual_seed(5); ctx2, h2, c2 = P.encoder_lstm(sd, toks, lens, bidirectional=bi, num_layers=nl, drop_ratio=0.5, drop=P.Drop("torch"))
        close(ctx, ctx2, "enc ctx train"); close(h, h2, "enc h train")
# envdrop decoder
dec = Pol.EnvDropDecoder(512, 0.5, 0.3).eval()
a = torch.randn(B,128); img = torch.randn(B,36,2176); cand = torch.randn(B,C,2176)
ht = torch.randn(B,512); c0 = torch.randn(B,512); ctx = torch.randn(B,L,512)
mask = torch.zeros(B,L,dtype=torch.bool); mask[1,10:] = True; mask[4,3:] = True
lo, (h1,c1), htl = dec(a, img.clone(), cand.clone(), ht, ht, c0, ctx, mask)
lo2, (h12,c12), htl2, _ = P.envdrop_decoder(dec.state_dict(), a, img, cand, ht, c0, ctx, mask)
close(lo, lo2, "envdrop logit", 1e-4); close(h1,h12,"h1"); close(c1,c12,"c1"); close(htl,htl2,"htilde")
dec.train(); torch.manual_seed(3); lo, (h1,c1), htl = dec(a, img.clone(), cand.clone(), ht, ht, c0, ctx, mask)
torch.manual_seed(3); lo2, (h12,c12), htl2, _ = P.envdrop_decoder(dec.state_dict(), a, img, cand, ht, c0, ctx, mask, drop=P.Drop("torch"))
close(lo, lo2, "envdrop logit train", 1e-3); close(htl,htl2,"htilde train")
# follower
dec = Pol.AttnDecoderLSTM(256, 0.5).eval()
ap = torch.randn(B,2176); h0 = torch.randn(B,256); c0 = torch.randn(B,256); ctx = torch.randn(B,L,256)
lo, (h1,c1), (ac, av) = dec(img, ap, cand, h0, c0, ctx, mask)
lo2, (h12,c12), (ac2, av2) = P.follower_decoder(dec.state_dict(), img, ap, cand, h0, c0, ctx, mask)
close(lo, lo2, "follower logit", 1e-4); close(h1,h12,"h1"); close(av,av2,"alpha_v"); close(ac,ac2,"alpha_c")
# monitor
for training in (False, True):
    dec = Pol.MonitorDecoder(512, 0.5, 80, [1024]); dec.train(training)
    if training:
        for m in dec.modules():
            if isinstance(m, torch.nn.Dropout): m.p = 0.0
    import copy
    sd = copy.deepcopy(dec.state_dict())
    h0 = torch.randn(B,512); c0 = torch.randn(B,512); ctx = torch.randn(B,80,512)
    m80 = torch.zeros(B,80,dtype=torch.bool); m80[:, 50:] = True
    cm = torch.zeros(B,C,dtype=torch.bool); cm[0,4:] = True; cm[2,2:] = True
    (lo, pr), (h1,c1), (ca, va) = dec(None, ap, cand, h0, c0, ctx, m80, cm)
    (lo2, pr2), (h12,c12), (ca2, va2) = P.monitor_decoder(sd, ap, cand, h0, c0, ctx, m80, cm, training=training)
    close(lo, lo2, f"monitor logit tr={training}", 2e-3); close(pr,pr2,"prog"); close(h1,h12,"h1",1e-4); close(va,va2,"cand attn",1e-4)
    if training:
        close(dec.state_dict()["proj_navigable_mlp.mlp.0.running_mean"], sd["proj_navigable_mlp.mlp.0.running_mean"], "bn rm")
        close(dec.state_dict()["proj_navigable_mlp.mlp.2.running_var"], sd["proj_navigable_mlp.mlp.2.running_var"], "bn rv", 1e-4)
cr = Pol.Critic(512, 0.5).eval(); close(cr(h0), P.critic(cr.state_dict(), h0), "critic")
print("OK")
