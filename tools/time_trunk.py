"""Context measurement for SURVEY.md section 8 row f2 (screening receptor reuse): the reference's once-per-complex trunk
`DiffusionConditioning` (PhysDock/models/layers/diffusion_conditioning.py:232-238: AtomEmbedder, Evoformer x4,
TemplatePairEmbedder, Pairformer x24; 102 M of the 153 M parameters) timed in PyTorch eager ON THE B200 with random
weights, next to the B200-native sampling of the same complex.  The UNMODIFIED reference package is imported from
baseline/_ref (pip --target install of /root/reference, git-ignored, travels with gpurun) through oracle/ref_import.py's
two shims; features are synthetic tensors of the shapes FeatureLoader produces.

    python tools/time_trunk.py [Nt Na]        (default 256 2048; BASELINE.json configs[3]: 8 samples per ligand, 40 steps)
"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PHYSDOCK_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))
from oracle.ref_import import import_reference, reference_available
from physdock_b200.synthetic import DiTDims, make_dit_state, token_layout


def synthetic_features(Nt, Na, n_msa=128, seed=0, device="cpu"):
    """A full FeatureLoader-shaped batch with random contents."""
    g = torch.Generator().manual_seed(seed)
    chunk, is_lig = token_layout(Nt, Na)
    a2t = torch.repeat_interleave(torch.arange(Nt), chunk)
    r = lambda *s: torch.randn(*s, generator=g)                                   # noqa: E731
    onehot = lambda n, c: torch.nn.functional.one_hot(torch.randint(0, c, (n,), generator=g), c).float()   # noqa: E731
    f = dict(
        atom_id_to_token_id=a2t, token_id_to_chunk_sizes=chunk, ap_mask=torch.ones(Na, Na), z_mask=torch.ones(Nt, Nt),
        a_mask=torch.ones(Na), is_ligand=is_lig.float(), x_gt=r(Na, 3) * 10, ref_pos=r(Na, 3) * 2,
        ref_feat=torch.cat([r(Na, 3), torch.zeros(Na, 1), onehot(Na, 128), torch.zeros(Na, 35)], -1),
        ref_space_uid=a2t.clone(), atom_id_to_conformer_id=a2t.clone(), target_feat=torch.cat([onehot(Nt, 32), torch.zeros(Nt, 33)], -1),
        msa_feat=r(n_msa, Nt, 34), rel_tok_feat=torch.zeros(Nt, Nt, 42), templ_feat=torch.cat([r(Nt, Nt, 39), torch.ones(Nt, Nt, 1)], -1),
        token_bonds=torch.zeros(Nt, Nt), token_bonds_feature=torch.zeros(Nt, Nt), key_res_feat=torch.zeros(Nt, 7),
        pocket_res_feat=torch.zeros(Nt), is_key_res=torch.zeros(Nt), is_protein=1 - is_lig.float(), is_dna=torch.zeros(Nt),
        is_rna=torch.zeros(Nt), s_mask=torch.ones(Nt), asym_id=is_lig.int(), sym_id=torch.zeros(Nt, dtype=torch.int32),
        entity_id=is_lig.int(), residue_index=torch.arange(Nt), restype=torch.randint(0, 20, (Nt,), generator=g),
        t_mask=torch.tensor(1.0))
    f["batch_msa_feat"] = f["msa_feat"][None]
    return {k: v.to(device) for k, v in f.items()}


def main():
    Nt, Na = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 2048)
    dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if not reference_available():
        raise SystemExit("baseline/_ref not found: python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>")
    PhysDock, PhysDockConfig, _, _ = import_reference()
    torch.manual_seed(0)
    model = PhysDock(PhysDockConfig(model_name="medium")).float().eval().to(dev)
    n_trunk = sum(p.numel() for p in model.diffusion_conditioning.parameters())
    batch = synthetic_features(Nt, Na, device=dev)
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    for tf32 in ((False, True) if dev.type == "cuda" else (False,)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.inference_mode():
            for _ in range(2):
                a, ap, s, z = model.diffusion_conditioning(batch)
            sync(); t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                a, ap, s, z = model.diffusion_conditioning(batch)
            sync(); ms = (time.perf_counter() - t0) / n * 1e3
        print(f"reference trunk (PyTorch eager, {dev.type}, TF32 {'on' if tf32 else 'off'}), Nt={Nt} Na={Na}, {n_trunk / 1e6:.1f} M params: {ms:9.1f} ms per complex")
    trunk_ms = ms
    if dev.type != "cuda":
        return
    # Is the eager trunk launch-bound?  Replay it from a CUDA graph (static shapes per complex) and compare.
    try:
        static = {k: v.clone() for k, v in batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.inference_mode():
            for _ in range(2):
                model.diffusion_conditioning(static)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.inference_mode(), torch.cuda.graph(g):
            out_g = model.diffusion_conditioning(static)
        g.replay(); sync(); t0 = time.perf_counter()
        for _ in range(3):
            g.replay()
        sync(); ms_g = (time.perf_counter() - t0) / 3 * 1e3
        err = max(float((x - y).abs().max()) for x, y in zip(out_g, (a, ap, s, z)))
        print(f"reference trunk replayed from a CUDA graph (TF32 on): {ms_g:9.1f} ms per complex (eager {trunk_ms:.1f}; max |diff| vs eager {err:.2e})")
    except Exception as e:  # noqa: BLE001
        print("reference trunk is not CUDA-graph capturable as written:", type(e).__name__, str(e)[:200])
        torch.cuda.synchronize()
    # the B200-native sampling of the same complex: 8 samples x 40 steps (BASELINE.json configs[3])
    from physdock_b200.dit import B200DiT
    from physdock_b200.sampler import sample_diffusion
    dims = DiTDims.named("medium")
    dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)
    a, ap, s, z = (t.float().clone() for t in (a, ap, s, z))
    for B in (8,):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            x = sample_diffusion(dit, batch, a, ap, s, z, num_sample=B, steps=40, karras_noise_schedule_power=1000, align_ref_pos=False)
            x_host = x.cpu()
            ms_s = (time.perf_counter() - t0) * 1e3
        print(f"B200 sampling of that complex, {B} samples x 40 steps (incl. pair-bias prepass + schedule conditioning): {ms_s:8.1f} ms per ligand")
        print(f"=> per screening ligand at Nt={Nt}/Na={Na}: trunk {trunk_ms:.0f} ms ({100 * trunk_ms / (trunk_ms + ms_s):.0f} %) + sampling {ms_s:.0f} ms; "
              f"1000 ligands / 8 GPUs = {(trunk_ms + ms_s) * 125 / 1e3:.0f} s of wall time per GPU")


if __name__ == "__main__":
    main()
