"""Host-side mirror of the reference's trainer/ package for the hot path: same function names, argument meaning and
error behaviour as edenartlab/sd-lora-trainer's trainer/{models,optimizer,loss,embedding_handler,
ti_cross_attn_loss}.py, dispatching into the B200 kernels."""
