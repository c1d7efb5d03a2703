mkdir -p gpurun_out
python tools/grad_noise.py --emulate > gpurun_out/r2i_grad_noise_emu.txt 2>&1
grep -v Warning gpurun_out/r2i_grad_noise_emu.txt | grep -v "Consider\|print(f"
