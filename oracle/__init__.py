"""CPU oracle for the EBSD dictionary-indexing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or, for the
CPU baseline, as the thing timed on the host cores) -- never as a fallback for
the CUDA path.

Each module restates one stage of the reference (poyentung/ebsd-vae, package
``latice``) and cites the reference file:line it follows:

* ``transform_ref``  -- 8-bit quantise + centre-crop        (latice/data_module.py:17-33)
* ``encoder_ref``    -- VAE encoder + mu/logvar heads       (latice/model.py:55-58, 93-129)
* ``topk_ref``       -- exact cosine top-k, canonical fp32  (latice/index/faiss_db.py:109-113, 216-256;
                        latice/index/chroma_db.py:231-259)
* ``consensus_ref``  -- quaternion consensus                (latice/index/chroma_db.py:261-375,
                        latice/index/faiss_db.py:258-393)

Pinning: ``make_golden.py`` (run in the build container, where ``/root/reference``
is mounted) imports the *unmodified* reference modules through ``refload`` and
writes small fixtures under ``tests/golden/``; ``tests/test_oracle_*.py`` check
every restatement against those fixtures.  The search stage is the exception:
the reference executes it inside chromadb/hnswlib or faiss, neither of which is
installable here, and its tests mock the call -- so the top-k oracle is
"parity unpinned" against the libraries and is anchored instead on an
independent float64 brute force (see ``topk_ref.py``).
"""
