"""ctc_pytorch_b200 — B200-native (sm_100a) CTC acoustic hot path behind the call surface of
Diamondfan/CTC_pytorch (CTC_Model / CTCLoss / GreedyDecoder / BeamDecoder). See DESIGN.md."""
__version__ = "0.1.0"
