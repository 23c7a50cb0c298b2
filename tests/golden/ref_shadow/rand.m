function r = rand(varargin)
% Deterministic stand-in for the interpreter's RNG while tests/golden/make_ref_vectors.m runs the UNMODIFIED reference.
% make_ref_vectors.m puts this directory first on the path, so every rand call of the reference lands here:
%   rand('seed', s)   (src/sparse_nmf.m:113)  -> switches to the Park-Miller "minimal standard" generator, state s
%   rand(m, n)        before any seed         -> the next m*n values of the committed stream tests/golden/rng_seed1.mat
%                     (src/init_buff.m:37-38: A_d = rand(R_d, m), then Ad_blk = rand(R_a, m_a))
%                     after a seed            -> Park-Miller values, column-major
% The same values reach the NumPy oracle through oracle.snmf_oracle.default_rng_inputs.
persistent mode state stream pos
if isempty(mode)
  mode = 0;
  pos = 0;
  here = fileparts(mfilename('fullpath'));
  s = load(fullfile(here, '..', 'rng_seed1.mat'));
  stream = s.stream(:);
end
if nargin >= 1 && ischar(varargin{1})
  if strcmp(varargin{1}, 'seed') && nargin == 2
    mode = 1;
    state = double(varargin{2});
    r = [];
    return;
  end
  error('rand shadow: unsupported call rand(''%s'', ...)', varargin{1});
end
if nargin == 0
  dims = [1 1];
elseif nargin == 1
  dims = double(varargin{1});
  if numel(dims) == 1, dims = [dims dims]; end
else
  dims = cellfun(@double, varargin);
end
n = prod(dims);
out = zeros(n, 1);
if mode == 0
  if pos + n > numel(stream), error('rand shadow: the committed stream is exhausted'); end
  out(:) = stream(pos + 1 : pos + n);
  pos = pos + n;
else
  for i = 1:n
    state = mod(16807 * state, 2147483647);
    out(i) = state / 2147483647;
  end
end
r = reshape(out, dims);
end
