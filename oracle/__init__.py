"""CPU oracle for the st-ito ES population-evaluation path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; nothing under ``st_ito_b200/`` does (tests/test_no_oracle_in_product.py
enforces it).  It restates, on the CPU, what the reference computes on the path
``run_es -> evaluate -> process_audio -> get_param_embeds -> Cnn14 -> cosine``:

================  ==========================================  =====================
module            follows (reference file:line)               parity status
================  ==========================================  =====================
dsp (EQ)          st_ito/effects.py:395-512, 784-873          pinned: golden vectors
                                                              made by executing the
                                                              reference's own code
dsp (chain walk)  st_ito/style_transfer.py:17-115, 324-359    pinned (same)
dsp (compressor,  st_ito/effects.py:876-959 -> pedalboard /   PARITY UNPINNED: JUCE
 reverb, dist,    JUCE, not vendored, no version pin          arithmetic restated
 delay)                                                       from recollection
frontend          st_ito/models/panns.py:139-168,219-245 ->   PARITY UNPINNED
                  torchlibrosa + librosa.filters.mel          (torchlibrosa absent);
                                                              cross-checked against
                                                              torch.stft/torchaudio
cnn14             st_ito/models/panns.py:25-80,180-281        pinned: golden vectors
                                                              from the reference's
                                                              Cnn14 body
embeds / fitness  st_ito/utils.py:444-508,                    restated; pinned via
                  st_ito/style_transfer.py:474-573            cnn14 goldens
================  ==========================================  =====================
"""
