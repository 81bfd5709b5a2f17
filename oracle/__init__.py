"""CPU oracle for the Semantic-NeRF hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``ucsa_neural_rendering_b200/`` may import this package.  The only
legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (where it is the
thing being *compared against*, never the thing shipped).

Contents
--------
``tcnn_spec``   frozen restatement of the tiny-cuda-nn arithmetic the reference
                calls (HashGrid, SphericalHarmonics deg 4, FullyFusedMLP).
                tiny-cuda-nn is an un-vendored, un-pinned dependency of the
                reference (README.md:51, git HEAD) and is absent from
                /root/reference, so this part is **parity unpinned**: it follows
                the published tcnn / Instant-NGP algorithm and is anchored on the
                reference's call sites (nr4seg/nerf/network_tcnn_semantics.py:36-100).
``live_path``   torch-CPU restatement of SemanticNeRFRenderer.run()/render() and
                sample_pdf (nr4seg/nerf/renderer_semantics.py:10-46,123-358) and of
                the network heads (network_tcnn_semantics.py:102-207).  **Pinned**:
                checked against the reference module itself imported in the build
                container (tests/golden/make_golden.py writes the fixtures).
``raymarch``    ctypes wrapper around ``raymarch_ref.c``, a plain-C restatement of
                the reference's CUDA ray-marching kernels
                (nr4seg/nerf/raymarching/src/raymarching.cu, pcg32.h).  Pinned on the
                GPU box against ``oracle/_ref`` (the reference .cu compiled as is).
"""
