"""Hand-built plaintext-HTTP segments for the Host / path-prefix stage:
(description, frame, expected name verdict).  Checked against the oracle on CPU and the GPU on B200."""
from clawker_b200.capture import frame, ip4, tls_client_hello

import name_cases as NC

RULES = NC.RULES + [
    {"dst": "api.example.com", "proto": "http", "port": 80, "action": "allow",
     "path_rules": [("/api/v1", "allow"), ("/admin", "deny")], "path_default": "deny"},            # envoy_test.go:271-296
    {"dst": ".datadoghq.com", "proto": "http"},                                                    # envoy_test.go:557-563
    {"dst": "claude.ai", "proto": "http"},
    {"dst": ".claude.ai", "proto": "http", "path_rules": [{"path": "/pub", "action": "Allow"}], "path_default": "allow"},
    {"dst": "open.example.net", "proto": "http", "port": 8080, "path_rules": [("/private", "deny")], "path_default": "ALLOW"},
    {"dst": "denied.example.net", "proto": "http", "action": "deny"},
]

ALLOW, DENY, NA = 1, 2, 0


def http(payload: bytes, dport: int = 80):
    return frame(NC.CT, ip4("93.184.216.34"), 6, 40002, dport, payload)


def req(host, path="/", method="GET", version="HTTP/1.1", extra=b"", host_line=None):
    hl = (b"Host: " + host.encode() + b"\r\n") if host_line is None else host_line
    return http(f"{method} {path} {version}\r\n".encode() + extra + hl + b"Accept: */*\r\n\r\n")


CASES = [
    ("wildcard rule covers the apex", req("datadoghq.com"), ALLOW),
    ("wildcard rule covers a subdomain", req("app.datadoghq.com"), ALLOW),
    ("deep subdomain", req("a.b.c.datadoghq.com", "/x/y?z=1"), ALLOW),
    ("host is matched case-insensitively", req("APP.DataDogHQ.com"), ALLOW),
    ("apex with a port: prefix wildcard d:*", req("datadoghq.com:8080"), ALLOW),
    ("subdomain with a port matches no form", req("app.datadoghq.com:8080"), DENY),
    ("suffix must start at a dot", req("notdatadoghq.com"), DENY),
    ("a wildcard never matches the empty string", req(".datadoghq.com"), DENY),
    ("prefix wildcard needs something after the colon", req("datadoghq.com:"), DENY),
    ("two colons: the shorter prefix still matches", req("datadoghq.com:80:90"), ALLOW),
    ("unknown host: deny_all", req("evil.example.com"), DENY),
    ("deny rule creates no virtual host", req("denied.example.net"), DENY),
    ("tls-only domain has no http virtual host", req("github.com"), DENY),
    ("path rule allow", req("api.example.com", "/api/v1/users"), ALLOW),
    ("prefix match is a string prefix", req("api.example.com", "/api/v10"), ALLOW),
    ("path rule with a query string", req("api.example.com", "/api/v1?x=1"), ALLOW),
    ("path rule deny", req("api.example.com", "/admin/panel"), DENY),
    ("path default deny", req("api.example.com", "/other"), DENY),
    ("path match is case sensitive", req("api.example.com", "/API/v1"), DENY),
    ("path shorter than the prefix", req("api.example.com", "/api"), DENY),
    ("path default allow", req("open.example.net", "/public", ), ALLOW),
    ("explicit deny before default allow", req("open.example.net", "/private/x"), DENY),
    ("exact rule owns the apex", req("claude.ai", "/zzz"), ALLOW),
    ("wildcard sibling: path rule", req("x.claude.ai", "/pub/1"), ALLOW),
    ("wildcard sibling: default allow", req("x.claude.ai", "/zzz"), ALLOW),
    ("POST", req("api.example.com", "/api/v1", method="POST"), ALLOW),
    ("HTTP/1.0 with Host", req("datadoghq.com", version="HTTP/1.0"), ALLOW),
    ("HTTP/1.0 without Host", http(b"GET / HTTP/1.0\r\nAccept: */*\r\n\r\n"), DENY),
    ("Host after other headers", req("datadoghq.com", extra=b"User-Agent: curl/8\r\nX-Long: " + b"a" * 200 + b"\r\n"), ALLOW),
    ("header name case and whitespace", req("", host_line=b"hOsT:  \t datadoghq.com \t \r\n"), ALLOW),
    ("empty Host value", req("", host_line=b"Host: \r\n"), DENY),
    ("first Host header wins", req("", host_line=b"Host: datadoghq.com\r\nHost: evil.example.com\r\n"), ALLOW),
    ("X-Host is not Host", req("", host_line=b"X-Host: datadoghq.com\r\nHost: evil.example.com\r\n"), DENY),
    ("absolute-form target: the authority wins", http(b"GET http://api.example.com/api/v1/x HTTP/1.1\r\nHost: evil.example.com\r\n\r\n"), ALLOW),
    ("absolute-form, scheme case", http(b"GET HTTP://datadoghq.com/ HTTP/1.1\r\nHost: evil.example.com\r\n\r\n"), ALLOW),
    ("absolute-form without a path is /", http(b"GET http://api.example.com HTTP/1.1\r\n\r\n"), DENY),
    ("absolute-form without a path, allow-all host", http(b"GET http://datadoghq.com HTTP/1.1\r\n\r\n"), ALLOW),
    ("absolute-form with an empty authority", http(b"GET http:///x HTTP/1.1\r\nHost: datadoghq.com\r\n\r\n"), DENY),
    ("OPTIONS * matches no / prefix", req("datadoghq.com", "*", method="OPTIONS"), DENY),
    ("CONNECT", http(b"CONNECT datadoghq.com:443 HTTP/1.1\r\nHost: datadoghq.com:443\r\n\r\n"), DENY),
    ("HTTP/2.0 in the request line", req("datadoghq.com", version="HTTP/2.0"), DENY),
    ("h2c preface", http(b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"), DENY),
    ("missing version", http(b"GET /\r\nHost: datadoghq.com\r\n\r\n"), DENY),
    ("extra token in the request line", http(b"GET / x HTTP/1.1\r\nHost: datadoghq.com\r\n\r\n"), DENY),
    ("binary garbage with a CRLF", http(b"\x00\x01\x02\r\n\r\n"), DENY),
    ("request line not complete", http(b"GET /a/very/long/path"), NA),
    ("bare LF line endings never complete a line", http(b"GET / HTTP/1.1\nHost: datadoghq.com\n\n"), NA),
    ("segment ends inside the header block, no Host yet", http(b"GET / HTTP/1.1\r\nUser-Agent: curl\r\nAccept: */"), NA),
    ("segment ends inside the header block after Host", http(b"GET / HTTP/1.1\r\nHost: datadoghq.com\r\nAccept: */"), ALLOW),
    ("Host line itself cut short", http(b"GET / HTTP/1.1\r\nHost: datadoghq.co"), NA),
    ("Host longer than 255", req("a" * 300 + ".datadoghq.com"), DENY),
    ("one CRLF only", http(b"\r\n"), DENY),
    ("port 8080 rule, request on 8080", frame(NC.CT, ip4("93.184.216.34"), 6, 40002, 8080, b"GET /public HTTP/1.1\r\nHost: open.example.net\r\n\r\n"), ALLOW),
    # TLS still takes the server_names path when HTTP rules exist
    ("TLS ClientHello next to HTTP rules", frame(NC.CT, ip4("93.184.216.34"), 6, 40001, 443, tls_client_hello("github.com")), ALLOW),
    ("handshake byte but not TLS", http(b"\x16 / HTTP/1.1\r\nHost: datadoghq.com\r\n\r\n"), DENY),
    ("empty TCP payload", http(b""), NA),
]
