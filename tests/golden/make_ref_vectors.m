% Reference vectors from the UNMODIFIED reference (GNU Octave or MATLAB): per-hop state of filewise_run_IS16.m on
% wav/M03_423C0213_STR.CH6.wav with the interpreter's RNG replaced by tests/golden/ref_shadow/rand.m.
%
%   cd <checkout of lordet01/SE_SNMF_NAT>
%   octave --no-gui --eval "repo='<this repo>'; run(fullfile(repo,'tests','golden','make_ref_vectors.m'))"
%
% writes <repo>/tests/golden/ref_vectors.mat (MAT v6); tests/test_oracle.py::test_oracle_matches_reference_vectors and
% tests/test_gpu_parity.py::test_gpu_matches_reference_vectors consume it when it exists.  Neither interpreter exists in
% the build image or on the GPU boxes of this project (profiles/r02_octave_probe.txt), so the file is not committed yet.
%
% The loop below is filewise_run_IS16.m:86-169 (same reads, same queueing, same overlap-add); init_buff.m,
% bnmf_sep_event_RT_IS16.m, sparse_nmf.m, blk_sparse.m, stft/istft helpers run from the reference's own src/.
if ~exist('repo', 'var'), error('set repo = path of the snmfnat-b200 checkout first'); end
addpath('src'); addpath('settings');
addpath(fullfile(repo, 'tests', 'golden', 'ref_shadow'), '-begin');   % rand.m shadow FIRST
clear rand;
initial_setting_SNMF_NAT;
fname = 'M03_423C0213_STR.CH6';
load(['basis/Clean_train_TIMIT_test/', 'TASLP_Splice0-SNMF_p2_DD0', '/R_100.mat']);
B_DFT_x = B_DFT_sub; B_Mel_x = B_Mel_sub;
load(['basis/CHiME3_bgn_ch6/', 'TASLP_Splice0-SNMF_p2_DD0', '/R_100.mat']);
B_DFT_d = B_DFT_sub; B_Mel_d = B_Mel_sub;
if strcmp(p.B_sep_mode, 'Mel')
  B1_x = B_Mel_x; B1_d = B_Mel_d;
else
  B1_x = B_DFT_x; B1_d = B_DFT_d;
end
g = init_buff(B1_x, B1_d, B_DFT_x, B_DFT_d, p);
fin = fopen(['wav/', fname, '.wav'], 'rb');
fread(fin, 22, 'int16');
frame_len = p.framelength; frame_shift = p.frameshift;
y = zeros(1, frame_len); x_tilde = zeros(1, frame_len);
n2 = p.fftlength / 2 + 1;
ref = struct();
ref.out_pcm = zeros(0, 1);
ref.Xm_tilde = zeros(n2, 0); ref.Ym = zeros(n2, 0); ref.A_d = zeros(p.R_d, 0); ref.lambda_dav = zeros(n2, 0);
ref.update_switch = zeros(1, 0);
l = 1; cnt_residue = 0;
while 1
  [~, len] = fread(fin, frame_shift, 'int16');
  if cnt_residue > p.delay, break; end
  if len ~= frame_shift
    cnt_residue = cnt_residue + 1;
    y = zeros(1, frame_len);
  else
    fseek(fin, -2 * frame_shift, 0);
    s_in = fread(fin, frame_shift, 'int16');
    y(1, 1:frame_len - frame_shift) = y(1, frame_shift + 1:frame_len);
    y(1, frame_len - frame_shift + 1:frame_len) = s_in;
  end
  if l <= p.init_N_len, g.W = 1; else, g.W = 0; end
  [~, ~, d_frame, g] = bnmf_sep_event_RT_IS16(y, l, g, p);
  ref.Xm_tilde(:, end + 1) = g.Xm_tilde(:, 1);
  ref.Ym(:, end + 1) = g.Ym(:, 1);
  ref.A_d(:, end + 1) = g.A_d(:, 1);
  ref.lambda_dav(:, end + 1) = g.lambda_dav(:, 1);
  ref.update_switch(end + 1) = g.update_switch;
  if l > p.delay
    x_tilde(1, 1:frame_len - frame_shift) = x_tilde(1, frame_shift + 1:frame_len);
    x_tilde(1, frame_len - frame_shift + 1:frame_len) = 0;
    x_tilde = x_tilde + d_frame(1, :);
    % fwrite(..., 'int16') of filewise_run_IS16.m:165 rounds to nearest and saturates
    ref.out_pcm = [ref.out_pcm; max(min(round(x_tilde(1, 1:frame_shift)'), 32767), -32768)];
  end
  l = l + 1;
end
fclose(fin);
ref.B_DFT_d_final = g.B_DFT_d;
ref.Ad_blk_final = g.Ad_blk;
ref.hops = l - 1;
outfile = fullfile(repo, 'tests', 'golden', 'ref_vectors.mat');
if exist('OCTAVE_VERSION', 'builtin')
  save('-v6', outfile, 'ref');
else
  save(outfile, 'ref', '-v6');
end
fprintf('wrote %s (%d hops)\n', outfile, ref.hops);
