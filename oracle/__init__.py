"""CPU oracle for the CLIP-guidance hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``clip_diffusion_b200`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU arm.

What is restated here, and what pins it
---------------------------------------
* ``oracle.resize_right``  -- restatement of the un-vendored third-party
  ``resize_right.resize`` (reference call sites: clip_diffusion/cutouts.py:6,64,105).
  No copy of that package exists under /root/reference => **parity unpinned**
  against the real package; cross-checked against torch's antialiased bicubic
  (interior pixels) and against identity/partition-of-unity properties.
* ``oracle.cutouts``       -- restatement of clip_diffusion/cutouts.py:47-134 that
  consumes an explicit RNG record.  Pinned **bit-exactly** against the
  reference's own ``Cutouts.forward`` executed in place (``oracle.ref_stubs``)
  in the build container; golden vectors committed under tests/golden/.
* ``oracle.losses``        -- restatement of clip_diffusion/losses.py:10-45, pinned
  against the reference module executed in place + golden vectors.
* ``oracle.clip_vit``      -- restatement of the un-vendored OpenAI CLIP
  VisionTransformer (reference call sites: utils/functional.py:101,
  models.py:76-80).  **Parity unpinned** against openai/CLIP itself; cross-checked
  against transformers' CLIPVisionModelWithProjection (quick_gelu) with mapped
  weights.
* ``oracle.cond_fn``       -- restatement of the closure sample.py:134-238.
* ``oracle.unet_norm``     -- the replicated UNet's normalisation blocks (guided-diffusion GroupNorm32 / scale-shift / SiLU, un-vendored:
  **parity unpinned** against the package; built at clip_diffusion/models.py:87-131) and the folded algebra csrc/unet_norm.cu implements,
  pinned against float64 autograd of the definition on the CPU.
"""
