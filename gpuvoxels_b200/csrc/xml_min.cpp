#include "xml_min.hpp"

#include <cctype>
#include <cstring>

namespace vx3b {

std::string xml_trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a]))
        a++;
    while (b > a && std::isspace((unsigned char)s[b - 1]))
        b--;
    return s.substr(a, b - a);
}

namespace {

struct Parser {
    const std::string &s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string &t) : s(t) {}

    bool starts(const char *lit) const { return s.compare(i, strlen(lit), lit) == 0; }
    void skip_ws() {
        while (i < s.size() && std::isspace((unsigned char)s[i]))
            i++;
    }
    static void append_decoded(std::string &out, const std::string &raw) {
        for (size_t k = 0; k < raw.size(); k++) {
            if (raw[k] != '&') {
                out += raw[k];
                continue;
            }
            size_t e = raw.find(';', k);
            if (e == std::string::npos) {
                out += raw[k];
                continue;
            }
            std::string ent = raw.substr(k + 1, e - k - 1);
            if (ent == "lt")
                out += '<';
            else if (ent == "gt")
                out += '>';
            else if (ent == "amp")
                out += '&';
            else if (ent == "quot")
                out += '"';
            else if (ent == "apos")
                out += '\'';
            else if (!ent.empty() && ent[0] == '#') {
                long code = ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X') ? strtol(ent.c_str() + 2, nullptr, 16) : strtol(ent.c_str() + 1, nullptr, 10);
                if (code > 0 && code < 128)
                    out += (char)code;
            } else {
                out += raw.substr(k, e - k + 1);
            }
            k = e;
        }
    }
    bool skip_misc() { // comments, PIs, doctype
        for (;;) {
            skip_ws();
            if (starts("<!--")) {
                size_t e = s.find("-->", i + 4);
                if (e == std::string::npos) {
                    err = "unterminated comment";
                    return false;
                }
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i + 2);
                if (e == std::string::npos) {
                    err = "unterminated processing instruction";
                    return false;
                }
                i = e + 2;
            } else if (starts("<!DOCTYPE")) {
                size_t e = s.find('>', i);
                if (e == std::string::npos) {
                    err = "unterminated doctype";
                    return false;
                }
                i = e + 1;
            } else
                return true;
        }
    }
    std::string name_tok() {
        size_t b = i;
        while (i < s.size() && !std::isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/' && s[i] != '=')
            i++;
        return s.substr(b, i - b);
    }
    std::unique_ptr<XmlNode> element() {
        // at '<'
        i++;
        auto n = std::make_unique<XmlNode>();
        n->name = name_tok();
        if (n->name.empty()) {
            err = "empty element name at offset " + std::to_string(i);
            return nullptr;
        }
        for (;;) { // attributes
            skip_ws();
            if (i >= s.size()) {
                err = "unexpected end inside <" + n->name + ">";
                return nullptr;
            }
            if (s[i] == '/') {
                if (i + 1 < s.size() && s[i + 1] == '>') {
                    i += 2;
                    return n;
                }
                err = "stray '/' in <" + n->name + ">";
                return nullptr;
            }
            if (s[i] == '>') {
                i++;
                break;
            }
            std::string an = name_tok();
            skip_ws();
            if (i >= s.size() || s[i] != '=') {
                err = "attribute without value in <" + n->name + ">";
                return nullptr;
            }
            i++;
            skip_ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) {
                err = "unquoted attribute in <" + n->name + ">";
                return nullptr;
            }
            char q = s[i++];
            size_t e = s.find(q, i);
            if (e == std::string::npos) {
                err = "unterminated attribute in <" + n->name + ">";
                return nullptr;
            }
            std::string v;
            append_decoded(v, s.substr(i, e - i));
            n->attrs.emplace_back(an, v);
            i = e + 1;
        }
        for (;;) { // content
            if (i >= s.size()) {
                err = "missing </" + n->name + ">";
                return nullptr;
            }
            if (s[i] != '<') {
                size_t e = s.find('<', i);
                if (e == std::string::npos)
                    e = s.size();
                append_decoded(n->text, s.substr(i, e - i));
                i = e;
                continue;
            }
            if (starts("<![CDATA[")) {
                size_t e = s.find("]]>", i + 9);
                if (e == std::string::npos) {
                    err = "unterminated CDATA in <" + n->name + ">";
                    return nullptr;
                }
                n->text.append(s, i + 9, e - (i + 9));
                i = e + 3;
                continue;
            }
            if (starts("<!--")) {
                size_t e = s.find("-->", i + 4);
                if (e == std::string::npos) {
                    err = "unterminated comment";
                    return nullptr;
                }
                i = e + 3;
                continue;
            }
            if (starts("<?")) {
                size_t e = s.find("?>", i + 2);
                if (e == std::string::npos) {
                    err = "unterminated processing instruction";
                    return nullptr;
                }
                i = e + 2;
                continue;
            }
            if (starts("</")) {
                i += 2;
                std::string cn = name_tok();
                skip_ws();
                if (i >= s.size() || s[i] != '>' || cn != n->name) {
                    err = "mismatched </" + cn + "> for <" + n->name + ">";
                    return nullptr;
                }
                i++;
                return n;
            }
            auto c = element();
            if (!c)
                return nullptr;
            n->children.push_back(std::move(c));
        }
    }
};

void escape_into(std::string &out, const std::string &s) {
    for (char c : s) {
        switch (c) {
        case '<':
            out += "&lt;";
            break;
        case '>':
            out += "&gt;";
            break;
        case '&':
            out += "&amp;";
            break;
        case '"':
            out += "&quot;";
            break;
        default:
            out += c;
        }
    }
}

void write_node(std::string &out, const XmlNode &n, int depth) {
    out.append(depth * 2, ' ');
    out += '<';
    out += n.name;
    for (auto &a : n.attrs) {
        out += ' ';
        out += a.first;
        out += "=\"";
        escape_into(out, a.second);
        out += '"';
    }
    std::string t = xml_trim(n.text);
    if (n.children.empty() && t.empty()) {
        out += "/>\n";
        return;
    }
    out += '>';
    if (!t.empty())
        escape_into(out, t);
    if (!n.children.empty()) {
        out += '\n';
        for (auto &c : n.children)
            write_node(out, *c, depth + 1);
        out.append(depth * 2, ' ');
    }
    out += "</";
    out += n.name;
    out += ">\n";
}

} // namespace

std::unique_ptr<XmlNode> xml_parse(const std::string &text, std::string *err) {
    Parser p(text);
    auto root = std::make_unique<XmlNode>();
    for (;;) {
        if (!p.skip_misc()) {
            if (err)
                *err = p.err;
            return nullptr;
        }
        if (p.i >= text.size())
            break;
        if (text[p.i] != '<') {
            if (err)
                *err = "text outside of an element at offset " + std::to_string(p.i);
            return nullptr;
        }
        auto e = p.element();
        if (!e) {
            if (err)
                *err = p.err;
            return nullptr;
        }
        root->children.push_back(std::move(e));
    }
    if (root->children.empty()) {
        if (err)
            *err = "no root element";
        return nullptr;
    }
    return root;
}

std::string xml_write(const XmlNode &node, bool declaration) {
    std::string out;
    if (declaration)
        out += "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n";
    write_node(out, node, 0);
    return out;
}

void xml_put_child(XmlNode &root, const std::string &path, std::unique_ptr<XmlNode> subtree) {
    XmlNode *cur = &root;
    size_t p = 0;
    for (;;) {
        size_t q = path.find('.', p);
        std::string part = path.substr(p, q == std::string::npos ? std::string::npos : q - p);
        bool last = q == std::string::npos;
        if (last) {
            subtree->name = part;
            for (auto &c : cur->children)
                if (c->name == part) {
                    c = std::move(subtree);
                    return;
                }
            cur->children.push_back(std::move(subtree));
            return;
        }
        XmlNode *nx = cur->child(part);
        if (!nx) {
            auto n = std::make_unique<XmlNode>();
            n->name = part;
            nx = n.get();
            cur->children.push_back(std::move(n));
        }
        cur = nx;
        p = q + 1;
    }
}

} // namespace vx3b
