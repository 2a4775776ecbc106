"""B200-native (sm_100a) RetinaNet loss / post-processing hot path.

Drop-in replacements for the reference's AnchorGenerator, SSD_loss, BBoxPredictor and nms
(NickTravers/NeuralNetworkLibrary: Applications/VisionModels/retinanet.py, Applications/Vision.py),
backed by hand-written CUDA kernels behind a C ABI (include/retina_b200.h).  There is no CPU
fallback: the wrappers raise if the CUDA library is missing.
"""
