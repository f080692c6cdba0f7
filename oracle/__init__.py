"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the ToucanTTS -> vocoder hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker (or as the timed
CPU baseline), never as the engine.

Contents
--------
``restate.py``          functional fp32 restatement (plain torch CPU ops) of the
                        reference hot path; every function cites the reference
                        file:line it follows.  Travels to the GPU box.
``factory.py``          deterministic state_dict / synthetic-input factory driven
                        by the committed key/shape manifests.
``alias_free_torch/``   restatement of the un-vendored third-party dependency
                        ``alias_free_torch~=0.0.6`` (reference requirements.txt)
                        that BigVGAN's AMP blocks import.
``shim.py``             import shim that makes the LIVE reference under
                        ``/root/reference`` importable (stub modules for the
                        missing audio/plot/phonemizer deps).  Only usable in the
                        authoring container; used to pin ``restate.py`` and to
                        generate ``tests/golden``.

Parity pin: the reference ships no golden vectors or known-answer tests
(SURVEY.md section 4), so ``restate.py`` is pinned against outputs of the
reference itself: live (``tests/test_oracle_vs_reference.py``, runs where
``/root/reference`` exists) and through the committed fixtures in
``tests/golden`` produced by ``oracle/make_golden.py``.
The ``alias_free_torch`` restatement has no reference-side pin at all
("parity unpinned" for that one dependency; cross-checked against the
independent copy of the same algorithm in transformers' qwen2_5_omni).
"""
