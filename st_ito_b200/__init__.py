"""st_ito_b200 -- B200-native implementation of st-ito's CMA-ES population-evaluation path.

Public surface (same names and argument meaning as the reference's st_ito package):
    st_ito_b200.style_transfer : run_es, process_audio, load_plugins, parameters_to_dict, savepop_to_disk
    st_ito_b200.effects        : Parameter, BasicParametricEQ, BasicCompressor, BasicDistortion, BasicDelay, BasicReverb
    st_ito_b200.utils          : load_param_model, get_param_embeds
    st_ito_b200.models.panns   : Cnn14
All arithmetic runs in libstito.so (hand-written sm_100a CUDA behind a C ABI, include/stito.h).
"""
__version__ = "0.1.0"
