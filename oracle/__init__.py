"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SiT hot path.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker.

Parity status:
* patch gather            -- pinned by the reference's own index tables (exact integers).
* SiT / MPP wrappers      -- pinned: ``oracle/sit_oracle.py`` is checked against the reference's own
                             ``models/sit.py`` / ``models/mpp.py`` imported from /root/reference (when present)
                             and against golden vectors generated from them (tests/golden).
* encoder arithmetic      -- PARITY UNPINNED by the reference itself: it lives in the third-party, un-pinned,
                             un-vendored ``vit-pytorch`` (requirements.txt:5).  ``oracle/vit_shim.py`` restates
                             the published pre-1.0 algorithm whose module layout is the one fixed by
                             ``utils/utils.py:18-33``; it is cross-checked against an independent
                             implementation (HuggingFace ``transformers`` ViT layers and torch SDPA).
"""
